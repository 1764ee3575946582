"""Pins the CPU oracle (oracle/ctc_oracle.c) -- no GPU needed.

1. against the known-answer vectors of the reference's own tests (tests/test_ctc.py:69-165,
   tests/test_ctc_decoder.py:44-166), stored in tests/golden/kat_*.npz;
2. against golden outputs produced by the reference itself (tests/golden/make_golden.py);
3. differentially against the compiled, unmodified reference engine when oracle/_ref/ is present.
"""
import os

import numpy as np
import pytest
import torch

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def T(a):
    return torch.from_numpy(np.array(a))   # keeps 0-dim arrays 0-dim


KATS = ["simple", "medium", "empty_label", "tf_1", "tf_2"]


@pytest.mark.parametrize("name", KATS)
def test_port_matches_reference_known_answers(name):
    g = gold("kat_loss")
    lp = T(g[name + "_lp"])
    eng = oracle.PortEngine(int(g[name + "_blank"]))
    losses, grads = eng.compute(lp, T(g[name + "_targets"]), T(g[name + "_ll"]), T(g[name + "_tl"]))
    assert abs(losses.sum().item() - float(g[name + "_expected"])) < 1e-5      # places=5 in the reference's test
    assert abs(losses.sum().item() - float(g[name + "_ref_loss"])) < 1e-6
    # the reference's leaf gradient (time-major module, after_logsoftmax=True) is the engine gradient
    ref_grad = T(g[name + "_ref_grad_tm"]).permute(1, 0, 2)
    assert torch.allclose(grads, ref_grad, rtol=1e-6, atol=1e-7, equal_nan=True)


@pytest.mark.parametrize("name", ["c1", "c2_b4", "c2_b4_peaky", "c4_b2", "edge_blank0", "edge_blank3"])
def test_port_engine_matches_golden(name):
    g = gold(name)
    lp = torch.log_softmax(T(g["x"]), 2)
    eng = oracle.PortEngine(int(g["blank"]))
    losses, grads = eng.compute(lp, T(g["targets"]), T(g["logits_lengths"]), T(g["targets_lengths"]))
    ref_l, ref_g = T(g["engine_losses"]), T(g["engine_grads"])
    assert torch.equal(torch.isinf(losses), torch.isinf(ref_l))
    assert torch.equal(torch.isnan(grads), torch.isnan(ref_g))
    assert torch.allclose(losses, ref_l, rtol=1e-6, atol=1e-6)
    assert torch.allclose(grads, ref_g, rtol=1e-6, atol=1e-7, equal_nan=True)


def test_c1_mean_loss_is_the_survey_value():
    g = gold("c1")
    assert abs(float(T(g["engine_losses"]).double().mean()) - 132.7469) < 1e-3


@pytest.mark.parametrize("name", ["c1", "edge_blank0", "edge_blank3"])
def test_module_restatement_matches_golden(name):
    """oracle.ctc_loss_module (modules/ctc_loss.py + functions/forward_backward.py restated) around
    the port engine reproduces the reference module's loss AND leaf gradient for every flag set."""
    g = gold(name)
    x = T(g["x"])
    i = 0
    while "m%d_flags" % i in g:
        reduce_, size_average, after, tm = [bool(v) for v in g["m%d_flags" % i]]
        xin = torch.log_softmax(x, 2) if after else x
        if tm:
            xin = xin.permute(1, 0, 2).contiguous()
        leaf = xin.clone().requires_grad_()
        loss = oracle.ctc_loss_module(oracle.PortEngine(int(g["blank"])), leaf, T(g["targets"]),
                                      T(g["logits_lengths"]), T(g["targets_lengths"]), reduce=reduce_ or None,
                                      size_average=size_average or None, after_logsoftmax=after, time_major=tm)
        (loss.sum() if loss.dim() else loss).backward()
        ref_l, ref_g = T(g["m%d_loss" % i]), T(g["m%d_grad" % i])
        assert loss.shape == ref_l.shape
        assert torch.allclose(loss.detach(), ref_l, rtol=1e-6, atol=1e-6, equal_nan=True)
        assert torch.equal(torch.isnan(leaf.grad), torch.isnan(ref_g))
        assert torch.allclose(leaf.grad, ref_g, rtol=1e-5, atol=1e-6, equal_nan=True)
        i += 1
    assert i > 0


def test_port_float64_matches_golden():
    g = gold("f64")
    leaf = T(g["x"]).clone().requires_grad_()
    loss = oracle.ctc_loss_module(oracle.PortEngine(0), leaf, T(g["targets"]), T(g["logits_lengths"]),
                                  T(g["targets_lengths"]))
    loss.sum().backward()
    assert torch.allclose(loss.detach(), T(g["m0_loss"]), rtol=1e-12, atol=1e-12)
    assert torch.allclose(leaf.grad, T(g["m0_grad"]), rtol=1e-10, atol=1e-12)
    # values recorded by the survey from the reference's own gradcheck fixture
    assert abs(loss[0].item() - 13.922620103088649) < 1e-9 and abs(loss[1].item() - 10.014536027275742) < 1e-9


def test_port_edge_semantics():
    """SURVEY 8(a) notes: infeasible -> +inf / all-NaN block incl. padding; padding rows = exp(lp);
    L=0 -> -sum lp[blank]; T=1,L=1 -> -lp[0][label]."""
    g = torch.Generator().manual_seed(3)
    lp = torch.log_softmax(torch.randn(4, 6, 5, generator=g), 2).double()
    tg = torch.tensor([[1, 1], [2, 0], [0, 0], [3, 4]])
    tl = torch.tensor([2, 1, 0, 2])
    ll = torch.tensor([2, 1, 4, 1])            # utt0 needs 3 frames (repeat) -> infeasible; utt3: T=1 < L=2
    losses, grads = oracle.PortEngine(0).compute(lp, tg, ll, tl)
    assert torch.isinf(losses[0]) and losses[0] > 0 and torch.isnan(grads[0]).all()
    assert torch.isinf(losses[3]) and torch.isnan(grads[3]).all()
    assert abs(losses[1].item() + lp[1, 0, 2].item()) < 1e-12
    assert abs(losses[2].item() + lp[2, :4, 0].sum().item()) < 1e-12
    assert torch.allclose(grads[1, 1:], lp[1, 1:].exp()) and torch.allclose(grads[2, 4:], lp[2, 4:].exp())
    with pytest.raises(ValueError):
        oracle.PortEngine(0).compute(lp, tg, torch.tensor([2, 0, 4, 1]), tl)    # T_i = 0 is UB in the reference


@pytest.mark.parametrize("name", ["simple", "sm", "probs_1", "probs_2"])
def test_port_greedy_known_answers(name):
    g = gold("kat_greedy")
    ll = g[name + "_ll"]
    lengths = None if ll[0] < 0 else T(ll)
    labels = [str(s) for s in g[name + "_labels"]]
    tg, ln, sents = oracle.greedy_decode(T(g[name + "_x"]), lengths, blank_idx=int(g[name + "_blank"]),
                                         labels=labels, prefer="port")
    assert sents == [str(s) for s in g[name + "_sentences"]]
    assert torch.equal(tg, T(g[name + "_targets"])) and torch.equal(ln, T(g[name + "_lengths"]))


def test_port_greedy_ties_nan_bf16():
    g = gold("greedy_random")
    x, ll, blank = T(g["x"]), T(g["ll"]), int(g["blank"])
    tg, ln, _ = oracle.greedy_decode(x, ll, blank_idx=blank, prefer="port")
    assert torch.equal(tg, T(g["targets"])) and torch.equal(ln, T(g["lengths"]))
    tg, ln, _ = oracle.greedy_decode(x, None, blank_idx=blank, prefer="port")
    assert torch.equal(tg, T(g["targets_full"])) and torch.equal(ln, T(g["lengths_full"]))
    xb = T(g["xb_bits"]).view(torch.bfloat16)
    tg, ln, _ = oracle.greedy_decode(xb, T(g["llb"]), blank_idx=0, prefer="port")
    assert torch.equal(tg, T(g["targets_b"])) and torch.equal(ln, T(g["lengths_b"]))


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("cfg,B", [("c1", 4), ("c2", 8), ("c3", 16), ("c4", 4), ("c5", 2)])
def test_port_vs_compiled_reference(cfg, B):
    """Differential: the C restatement against the unmodified reference engine on the seeded
    synthetic draws of every BASELINE config (sub-batches)."""
    _, T_, V, Lmin, Lmax, seed, _, full = oracle.CONFIGS[cfg]
    x, tg, ll, tl = oracle.make_inputs(B, T_, V, Lmin, Lmax, seed, full_length=full)
    lp = torch.log_softmax(x, 2)
    l1, g1 = oracle.PortEngine(0).compute(lp, tg, ll, tl)
    l2, g2 = oracle.ref_engine(0).compute(lp, tg, ll, tl)
    assert torch.equal(l1, l2) or torch.allclose(l1, l2, rtol=1e-7, atol=0)
    assert torch.allclose(g1, g2, rtol=0, atol=1e-7)


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_greedy_vs_compiled_reference():
    x, _, ll, _ = oracle.make_inputs(16, 250, 1024, 40, 80, 3)
    a = oracle.greedy_decode(x, ll, prefer="port")
    b = oracle.greedy_decode(x, ll, prefer="reference")
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_log_softmax_restatement():
    x = torch.randn(50, 29, generator=torch.Generator().manual_seed(0))
    out = torch.from_numpy(oracle.log_softmax_f32(x.numpy()))
    assert torch.allclose(out, torch.log_softmax(x, -1), rtol=0, atol=1e-6)


# --------------------------------------------------------------------------------------------
# Viterbi forced alignment: the C restatement against vectors the reference's own
# numba code produced (tests/golden/make_align_golden.py)
# --------------------------------------------------------------------------------------------
ALIGN_GOLDENS = ["align_c1", "align_c2_b8", "align_c2_b4_peaky", "align_c4_b2", "align_c5_b2", "align_asg_c1",
                 "align_asg_c2_b4", "align_edge", "align_asg_edge"]


@pytest.mark.parametrize("name", ALIGN_GOLDENS)
def test_alignment_oracle_matches_reference_goldens(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    out = oracle.get_alignment_3d(torch.from_numpy(g["log_probs"]), torch.from_numpy(g["targets"]),
                                  torch.from_numpy(g["logits_lengths"]), torch.from_numpy(g["targets_lengths"]),
                                  is_ctc=bool(g["is_ctc"]))
    assert out.dtype == torch.int64 and torch.equal(out, torch.from_numpy(g["aligned"]))


# --------------------------------------------------------------------------------------------
# CTC without blank: the C restatement against the reference's own numba code
# --------------------------------------------------------------------------------------------
NOBLANK_GOLDENS = ["noblank_c1", "noblank_c1_space", "noblank_c2_b4", "noblank_c2_b4_space", "noblank_edge_space", "noblank_edge"]


@pytest.mark.parametrize("name", NOBLANK_GOLDENS)
def test_noblank_oracle_matches_reference_goldens(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    losses, grads = oracle.ctc_without_blank(torch.from_numpy(g["log_probs"]), torch.from_numpy(g["targets"]),
                                             torch.from_numpy(g["logits_lengths"]), torch.from_numpy(g["targets_lengths"]),
                                             space_idx=int(g["space_idx"]))
    rl, rg = torch.from_numpy(g["losses"]), torch.from_numpy(g["grads"]).double()
    assert torch.equal(torch.isnan(grads), torch.isnan(rg)) and torch.equal(torch.isinf(losses), torch.isinf(rl))
    fin = torch.isfinite(rl)
    torch.testing.assert_close(losses[fin], rl[fin], rtol=1e-12, atol=1e-12)
    ok = ~torch.isnan(rg)
    # the reference's grads array has the dtype of its float32 input (np.zeros_like(logits)): one float32 rounding
    torch.testing.assert_close(grads[ok], rg[ok], rtol=0, atol=2e-7)


# --------------------------------------------------------------------------------------------
# LM-free prefix beam search: the C restatement against the compiled reference
# --------------------------------------------------------------------------------------------
BEAM_GOLDENS = sorted(os.path.basename(p)[:-4] for p in __import__("glob").glob(os.path.join(GOLD, "beam_*.npz")))


@pytest.mark.parametrize("name", BEAM_GOLDENS)
def test_beam_oracle_matches_reference_goldens(name):
    """tests/golden/beam_*.npz hold what the reference's compiled decoder returned (make_beam_golden.py).  Every
    utterance whose prunes never had equal scores on both sides of the cut must be reproduced symbol for symbol."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    labels = str(g["labels"]).split("\x1f") if str(g["labels"]) else []
    tgt, tl, sents, ties = oracle.beam_decode(torch.from_numpy(g["log_probs"]), torch.from_numpy(g["logits_lengths"]),
                                              blank_idx=int(g["blank"]), beam_width=int(g["beam"]), labels=labels,
                                              after_logsoftmax=True, wip=float(g["wip"]), prefer="port", return_ties=True)
    assert torch.equal(ties, torch.from_numpy(g["ties"]))
    want_t, want_l = torch.from_numpy(g["targets"]), torch.from_numpy(g["targets_lengths"])
    n_free = 0
    for i in range(tl.numel()):
        if ties[i] == 0:
            n_free += 1
            assert int(tl[i]) == int(want_l[i]) and tgt[i, :int(tl[i])].tolist() == want_t[i, :int(tl[i])].tolist(), i
    assert n_free > 0 or "ties" in name
    if n_free == tl.numel() and labels and int(want_t.min()) >= 0:
        assert sents == str(g["sentences"]).split("\x1f")


def test_beam_oracle_known_answers():
    """The reference's own beam-search known answers (tests/test_ctc_decoder.py:86-166)."""
    for name, want in (("beam_kat_sm", "a"), ("beam_kat_1", "acdc"), ("beam_kat_2", "b'a")):
        g = np.load(os.path.join(GOLD, name + ".npz"))
        labels = str(g["labels"]).split("\x1f")
        out = oracle.beam_decode(torch.from_numpy(g["log_probs"]), torch.from_numpy(g["logits_lengths"]), blank_idx=int(g["blank"]),
                                 beam_width=20, labels=labels, after_logsoftmax=True, wip=0.0, prefer="port")
        assert out[2] == [want]


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
def test_beam_oracle_differential_vs_compiled_reference():
    """Random shapes: wherever the restatement reports no tie it equals the compiled reference."""
    g = torch.Generator().manual_seed(123)
    free = 0
    for V in (3, 6, 29):
        for beam in (2, 9, 100):
            for T in (3, 25, 60):
                lp = torch.log_softmax(torch.randn(2, T, V, generator=g) * 2.0, 2)
                ll = torch.tensor([T, max(1, T - 2)])
                a = oracle.beam_decode(lp, ll, beam_width=beam, after_logsoftmax=True, prefer="reference")
                b = oracle.beam_decode(lp, ll, beam_width=beam, after_logsoftmax=True, prefer="port", return_ties=True)
                for i in range(2):
                    if b[3][i] == 0:
                        free += 1
                        assert int(a[1][i]) == int(b[1][i]) and a[0][i, :int(a[1][i])].tolist() == b[0][i, :int(b[1][i])].tolist()
    assert free >= 40
