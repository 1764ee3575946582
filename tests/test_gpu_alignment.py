"""GPU parity of the Viterbi forced alignment: csrc/ctc_viterbi.cu through
end2end_b200.utils.alignment.get_alignment_3d against (a) vectors produced by the reference's own numba code
(tests/golden/align_*.npz, made by make_align_golden.py), (b) the C oracle on the BASELINE shapes.  Bit-exact."""
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDENS = ["align_c1", "align_c2_b8", "align_c2_b4_peaky", "align_c4_b2", "align_c5_b2", "align_asg_c1",
           "align_asg_c2_b4", "align_edge", "align_asg_edge"]


def T(a):
    return torch.from_numpy(np.array(a))


@pytest.mark.parametrize("name", GOLDENS)
def test_alignment_matches_reference_goldens(name):
    from end2end_b200.utils.alignment import get_alignment_3d
    g = np.load(os.path.join(GOLD, name + ".npz"))
    lp, tg, ll, tl = T(g["log_probs"]), T(g["targets"]), T(g["logits_lengths"]), T(g["targets_lengths"])
    out = get_alignment_3d(lp, tg, ll, tl, is_ctc=bool(g["is_ctc"]))              # CPU tensors in
    assert out.dtype == torch.int64 and not out.is_cuda and torch.equal(out, T(g["aligned"]))
    out2 = get_alignment_3d(lp.cuda(), tg.cuda(), ll.cuda().int(), tl.cuda().int(), is_ctc=bool(g["is_ctc"]))   # CUDA, int32 lengths
    assert torch.equal(out2, T(g["aligned"]))


@pytest.mark.parametrize("cfg,B", [("c1", 4), ("c2", 64), ("c3", 256), ("c4", 32), ("c5", 16)])
@pytest.mark.parametrize("is_ctc", [True, False])
def test_alignment_baseline_shapes_vs_oracle(cfg, B, is_ctc):
    from end2end_b200.utils.alignment import get_alignment_3d_device
    _, T_, V, Lmin, Lmax, seed, dtype, full = oracle.CONFIGS[cfg]
    x, tg, ll, tl = oracle.make_inputs(B, T_, V, Lmin, Lmax, seed, dtype=torch.float32, full_length=full)
    lp = torch.log_softmax(x, 2)
    ref = oracle.get_alignment_3d(lp, tg, ll, tl, is_ctc=is_ctc)
    out = get_alignment_3d_device(lp.cuda(), tg.cuda(), ll.cuda(), tl.cuda(), is_ctc=is_ctc)
    assert out.is_cuda and torch.equal(out.cpu(), ref)
    # peaky variant (x5): long runs of one label, many near-ties in the tails
    lp5 = torch.log_softmax(x * 5, 2)
    assert torch.equal(get_alignment_3d_device(lp5.cuda(), tg.cuda(), ll.cuda(), tl.cuda(), is_ctc=is_ctc).cpu(),
                       oracle.get_alignment_3d(lp5, tg, ll, tl, is_ctc=is_ctc))


def test_alignment_properties_full_size_c5():
    """BASELINE config 5 at a batch the oracle cannot finish quickly: every aligned row collapses (repeats merged,
    blanks dropped) to exactly its targets; frames past the utterance are -100."""
    from end2end_b200.utils.alignment import get_alignment_3d_device
    B, T_, V, Lmin, Lmax, seed = 512, 1600, 29, 300, 600, 4
    x, tg, ll, tl = oracle.make_inputs(B, T_, V, Lmin, Lmax, seed, dtype=torch.float32, full_length=False)
    out = get_alignment_3d_device(torch.log_softmax(x.cuda(), 2), tg.cuda(), ll.cuda(), tl.cuda()).cpu()
    for b in range(0, B, 37):
        n, L = int(ll[b]), int(tl[b])
        row = out[b, :n]
        assert bool((out[b, n:] == -100).all())
        keep = torch.ones(n, dtype=torch.bool)
        keep[1:] = row[1:] != row[:-1]
        collapsed = row[keep & (row != 0)]
        # a label repeated in the targets is separated by a blank on any valid path, so collapsing recovers the targets
        assert torch.equal(collapsed, tg[b, :L]), b


def test_alignment_time_major_strides_blank_and_errors():
    from end2end_b200.utils.alignment import get_alignment_3d_device
    x, tg, ll, tl = oracle.make_inputs(6, 40, 9, 2, 9, 3)
    lp = torch.log_softmax(x, 2)
    ref = oracle.get_alignment_3d(lp, tg, ll, tl)
    tm = lp.permute(1, 0, 2).contiguous().cuda().permute(1, 0, 2)          # time-major storage, batch-major view: read in place
    assert torch.equal(get_alignment_3d_device(tm, tg.cuda(), ll.cuda(), tl.cuda()).cpu(), ref)
    for dt in (torch.float64, torch.bfloat16):
        lpd = lp.to(dt)
        assert torch.equal(get_alignment_3d_device(lpd.cuda(), tg.cuda(), ll.cuda(), tl.cuda()).cpu(),
                           oracle.get_alignment_3d(lpd.double(), tg, ll, tl))
    # another blank index: the extended targets use it
    tg4 = torch.where(tg == 4, torch.zeros_like(tg), tg)
    assert torch.equal(get_alignment_3d_device(lp.cuda(), tg4.cuda(), ll.cuda(), tl.cuda(), blank_idx=4).cpu(),
                       oracle.get_alignment_3d(lp, tg4, ll, tl, blank_idx=4))
    with pytest.raises(ValueError):
        get_alignment_3d_device(lp.cuda(), tg[:3].cuda(), ll.cuda(), tl.cuda())
    with pytest.raises(ValueError):
        get_alignment_3d_device(lp[0].cuda(), tg.cuda(), ll.cuda(), tl.cuda())


def test_aligned_targets_loss_module_golden():
    """The consumer module (reference modules/alignment_loss.py:7-33) on CPU and CUDA tensors."""
    from end2end_b200.modules.alignment_loss import AlignedTargetsLoss
    import pytorch_end2end.modules.alignment_loss as alias
    assert alias.AlignedTargetsLoss is AlignedTargetsLoss
    g = np.load(os.path.join(GOLD, "align_loss_module.npz"))
    lp, tg, ll, tl = T(g["log_probs"]), T(g["targets"]), T(g["logits_lengths"]), T(g["targets_lengths"])
    for is_ctc in (True, False):
        for ib in (False, True):
            want = T(g["loss_ctc%d_ib%d" % (is_ctc, ib)])
            got = AlignedTargetsLoss(is_ctc, ignore_blank=ib)(lp, tg, ll, tl)
            torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)
            leaf = lp.cuda().requires_grad_()
            got_c = AlignedTargetsLoss(is_ctc, ignore_blank=ib)(leaf, tg.cuda(), ll.cuda(), tl.cuda())
            torch.testing.assert_close(got_c.cpu(), want, rtol=1e-6, atol=1e-6)
            got_c.sum().backward()                                   # NLL on the alignment is differentiable in log_probs
            assert leaf.grad is not None and bool(torch.isfinite(leaf.grad).all())
