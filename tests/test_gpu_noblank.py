"""GPU parity of CTC without blank: csrc/ctc_noblank.cu through the reference's Function / module
names against (a) vectors produced by the reference's own numba code (tests/golden/noblank_*.npz, made by
make_noblank_golden.py), (b) the C oracle on larger batches.  Tolerance: rel 1e-5 / abs 1e-5 (north_star), NaN / inf
positions identical."""
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDENS = ["noblank_c1", "noblank_c1_space", "noblank_c2_b4", "noblank_c2_b4_space", "noblank_edge_space", "noblank_edge"]
RTOL, ATOL = 1e-5, 1e-5


def T(a):
    return torch.from_numpy(np.array(a))


def close(ours, ref, what):
    ours, ref = ours.detach().cpu().double(), ref.detach().cpu().double()
    assert ours.shape == ref.shape, what
    assert torch.equal(torch.isnan(ours), torch.isnan(ref)), what + ": NaN positions differ"
    assert torch.equal(torch.isinf(ours), torch.isinf(ref)), what + ": inf positions differ"
    fin = torch.isfinite(ref)
    err = (ours[fin] - ref[fin]).abs()
    assert bool((err <= ATOL + RTOL * ref[fin].abs()).all()), "%s: max err %.3e" % (what, float(err.max()))


@pytest.mark.parametrize("name", GOLDENS)
def test_noblank_matches_reference_goldens(name):
    from end2end_b200.functions.ctc_without_blank import ctc_without_blank_3d_loss
    g = np.load(os.path.join(GOLD, name + ".npz"))
    lp, tg, ll, tl = T(g["log_probs"]), T(g["targets"]), T(g["logits_lengths"]), T(g["targets_lengths"])
    losses, grads = ctc_without_blank_3d_loss(lp.cuda(), tg.cuda(), ll.cuda(), tl.cuda(), int(g["space_idx"]))
    assert losses.dtype == torch.float32 and grads.dtype == torch.float32 and grads.is_cuda
    close(losses, T(g["losses"]), name + " losses")
    close(grads, T(g["grads"]), name + " grads")
    l2, g2 = ctc_without_blank_3d_loss(lp, tg.int(), ll.int(), tl.int(), int(g["space_idx"]))      # CPU tensors, int32 indices
    assert torch.equal(l2.cpu(), losses.cpu()) and torch.equal(torch.nan_to_num(g2.cpu()), torch.nan_to_num(grads.cpu()))


@pytest.mark.parametrize("cfg,B,space", [("c1", 4, -1), ("c2", 64, -1), ("c2", 64, 5), ("c3", 128, 0), ("c4", 16, -1), ("c5", 4, 7)])
def test_noblank_baseline_shapes_vs_oracle(cfg, B, space):
    from end2end_b200.functions.ctc_without_blank import ctc_without_blank_3d_loss
    _, T_, V, Lmin, Lmax, seed, dtype, full = oracle.CONFIGS[cfg]
    x, tg, ll, tl = oracle.make_inputs(B, T_, V, Lmin, Lmax, seed, dtype=torch.float32, full_length=full)
    lp = torch.log_softmax(x, 2)
    l_ref, g_ref = oracle.ctc_without_blank(lp, tg, ll, tl, space)
    losses, grads = ctc_without_blank_3d_loss(lp.cuda(), tg.cuda(), ll.cuda(), tl.cuda(), space)
    close(losses, l_ref, cfg + " losses")
    close(grads, g_ref, cfg + " grads")
    # float64 input: computed and returned in float64
    l64, g64 = ctc_without_blank_3d_loss(lp[:2].double().cuda(), tg[:2].cuda(), ll[:2].cuda(), tl[:2].cuda(), space)
    assert l64.dtype == torch.float64
    close(l64, l_ref[:2], cfg + " f64 losses")


def test_noblank_module_golden_and_gradcheck():
    from end2end_b200.modules.ctc_without_blank import CTCWithoutBlankLoss
    import pytorch_end2end.modules.ctc_without_blank as alias
    assert alias.CTCWithoutBlankLoss is CTCWithoutBlankLoss
    g = np.load(os.path.join(GOLD, "noblank_module.npz"))
    logits, tg, ll, tl = T(g["logits"]), T(g["targets"]), T(g["logits_lengths"]), T(g["targets_lengths"])
    for space in (-1, 3):
        for reduce in (True, False):
            for dev in ("cuda", "cpu"):
                leaf = logits.detach().clone().to(dev).requires_grad_()
                loss = CTCWithoutBlankLoss(reduce=reduce, space_idx=space)(leaf, tg.to(dev), ll.to(dev), tl.to(dev))
                assert loss.device.type == dev
                loss.sum().backward()
                close(loss, T(g["loss_s%d_r%d" % (space, reduce)]), "module loss")
                close(leaf.grad, T(g["grad_s%d_r%d" % (space, reduce)]), "module grad")
    # probabilities in (after_softmax=True) give the same loss
    probs = torch.softmax(logits, 2).cuda()
    a = CTCWithoutBlankLoss(reduce=False, after_softmax=True)(probs, tg.cuda(), ll.cuda(), tl.cuda())
    close(a, T(g["loss_s-1_r0"]), "after_softmax loss")
    # analytic vs numerical gradient in float64, through the module: the Function's gradient is exp(lp) - posterior (the
    # softmax Jacobian folded in, as the reference returns it), which is the true gradient only behind the log_softmax
    x = torch.randn(2, 12, 5, dtype=torch.float64, generator=torch.Generator().manual_seed(1)).cuda().requires_grad_()
    tg2, ll2, tl2 = torch.tensor([[1, 2, 2], [4, 0, 1]]).cuda(), torch.tensor([12, 9]).cuda(), torch.tensor([3, 2]).cuda()
    crit = CTCWithoutBlankLoss(reduce=False)
    assert torch.autograd.gradcheck(lambda z: crit(z, tg2, ll2, tl2), (x,), eps=1e-6, atol=1e-4)


def test_noblank_invalid_arguments():
    from end2end_b200.functions.ctc_without_blank import ctc_without_blank_3d_loss
    lp = torch.log_softmax(torch.randn(2, 6, 4), 2).cuda()
    tg, ll, tl = torch.tensor([[1, 2], [3, 1]]).cuda(), torch.tensor([6, 6]).cuda(), torch.tensor([2, 2]).cuda()
    with pytest.raises(ValueError):
        ctc_without_blank_3d_loss(lp, tg, ll, tl, 4)                        # space_idx outside the alphabet
    with pytest.raises(ValueError):
        ctc_without_blank_3d_loss(lp[0], tg, ll, tl)
    losses, grads = ctc_without_blank_3d_loss(lp, torch.tensor([[1, 9], [3, 1]]).cuda(), ll, tl)     # a label outside the alphabet
    assert torch.isnan(losses[0]) and torch.isnan(grads[0]).all() and torch.isfinite(losses[1])
