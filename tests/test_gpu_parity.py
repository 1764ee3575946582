"""GPU parity tests (run on the B200 box: pytest -m gpu).

The CUDA path (through the public modules, the engine, and the raw C ABI) is compared with
* the golden vectors the reference itself produced (tests/golden, made by make_golden.py),
* the CPU oracle (the compiled reference when oracle/_ref travelled, else the C port) on the
  seeded synthetic draws of every BASELINE config at sizes the oracle finishes in seconds,
* size-independent properties at BASELINE's full sizes.

Tolerances (BASELINE.json north_star): loss and logits-gradient rel 1e-5 / abs 1e-5 in fp32
(NaN / inf positions must coincide); bf16 logits: fp32-accurate loss (1e-5) and gradients within
one bf16 ulp (rel 2^-8, abs 1e-5) of the reference run on logits.float(); greedy decodes bit-exact.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL, ATOL = 1e-5, 1e-5
BF16_RTOL = 2.0 ** -8


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def T(a):
    return torch.from_numpy(np.array(a))


def assert_parity(ours, ref, rtol=RTOL, atol=ATOL, what=""):
    ours, ref = ours.detach().cpu().double(), ref.detach().cpu().double()
    assert ours.shape == ref.shape, (what, ours.shape, ref.shape)
    assert torch.equal(torch.isnan(ours), torch.isnan(ref)), what + ": NaN positions differ"
    assert torch.equal(torch.isposinf(ours), torch.isposinf(ref)), what + ": +inf positions differ"
    fin = torch.isfinite(ref)
    err = (ours[fin] - ref[fin]).abs()
    bad = err > atol + rtol * ref[fin].abs()
    assert not bool(bad.any()), "%s: %d/%d outside tolerance, max abs err %.3e" % (
        what, int(bad.sum()), err.numel(), float(err.max()))


@pytest.fixture(scope="module")
def e2e():
    import end2end_b200
    from end2end_b200 import _lib
    _lib.load()
    return end2end_b200


def cuda(*ts):
    return [t.cuda() if t is not None else None for t in ts]


# --------------------------------------------------------------------------------------------
# the reference's own known-answer tests, driven the way tests/test_ctc.py:22-66 drives them
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["simple", "medium", "empty_label", "tf_1", "tf_2"])
def test_reference_known_answers(e2e, name):
    g = gold("kat_loss")
    lp_tm = T(g[name + "_lp"]).permute(1, 0, 2).contiguous()            # time major, as run_grads does
    tg, ll, tl = T(g[name + "_targets"]), T(g[name + "_ll"]), T(g[name + "_tl"])
    crit = e2e.CTCLoss(reduce=True, size_average=False, after_logsoftmax=True, time_major=True,
                       blank_idx=int(g[name + "_blank"]))
    costs = []
    for dev in ("cpu", "cuda"):                                        # host-tensor path and device path
        leaf = lp_tm.detach().clone().to(dev).requires_grad_()
        cost = crit(leaf, tg, ll, tl)                                   # int32 CPU targets/lengths, as in the reference
        cost.backward()
        costs.append(cost.item())
        assert abs(cost.item() - float(g[name + "_expected"])) < 1e-5   # assertAlmostEqual(places=5)
        assert_parity(leaf.grad, T(g[name + "_ref_grad_tm"]), what=name + " grad " + dev)
    assert abs(costs[0] - costs[1]) < 1e-5                              # cpu_cost == gpu_cost


# --------------------------------------------------------------------------------------------
# golden vectors produced by the reference (engine contract + module flag sets)
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1", "c2_b4", "c2_b4_peaky", "c4_b2", "edge_blank0", "edge_blank3"])
def test_engine_contract_golden(e2e, name):
    g = gold(name)
    lp = torch.log_softmax(T(g["x"]), 2)
    tg, ll, tl = T(g["targets"]), T(g["logits_lengths"]), T(g["targets_lengths"])
    eng = e2e.CTCLossEngine(int(g["blank"]))
    for tensors in ((lp.cuda(), *cuda(tg, ll, tl)), (lp.cuda(), tg, ll, tl), (lp, tg, ll, tl)):
        losses, grads = eng.compute(*tensors)
        assert losses.device == tensors[0].device and grads.device == tensors[0].device
        assert losses.dtype == lp.dtype and grads.shape == lp.shape
        assert_parity(losses, T(g["engine_losses"]), what=name + " losses")
        assert_parity(grads, T(g["engine_grads"]), what=name + " grads")   # incl. padding rows = exp(lp), NaN blocks


@pytest.mark.parametrize("name", ["c1", "c2_b4", "c2_b4_peaky", "c4_b2", "edge_blank0", "edge_blank3"])
def test_module_golden(e2e, name):
    g = gold(name)
    x = T(g["x"])
    tg, ll, tl = cuda(T(g["targets"]), T(g["logits_lengths"]), T(g["targets_lengths"]))
    i = 0
    while "m%d_flags" % i in g:
        reduce_, size_average, after, tm = [bool(v) for v in g["m%d_flags" % i]]
        xin = torch.log_softmax(x, 2) if after else x
        if tm:
            xin = xin.permute(1, 0, 2).contiguous()
        leaf = xin.cuda().requires_grad_()
        crit = e2e.CTCLoss(reduce=reduce_ or None, size_average=size_average or None, after_logsoftmax=after,
                           time_major=tm, blank_idx=int(g["blank"]))
        loss = crit(leaf, tg, ll, tl)
        (loss.sum() if loss.dim() else loss).backward()
        assert_parity(loss, T(g["m%d_loss" % i]), what="%s mode %d loss" % (name, i))
        assert_parity(leaf.grad, T(g["m%d_grad" % i]), what="%s mode %d grad" % (name, i))
        assert leaf.grad.is_contiguous()
        i += 1
    assert i > 0


def test_float64_golden_and_gradcheck(e2e):
    g = gold("f64")
    x, tg, ll, tl = T(g["x"]), T(g["targets"]), T(g["logits_lengths"]), T(g["targets_lengths"])
    leaf = x.cuda().requires_grad_()
    crit = e2e.CTCLoss(blank_idx=0, time_major=False, after_logsoftmax=False)
    loss = crit(leaf, *cuda(tg, ll, tl))
    loss.sum().backward()
    assert loss.dtype == torch.float64
    assert_parity(loss, T(g["m0_loss"]), rtol=1e-9, atol=1e-9, what="f64 loss")
    assert_parity(leaf.grad, T(g["m0_grad"]), rtol=1e-6, atol=1e-6, what="f64 grad")
    # the reference's test_gradient (tests/test_ctc.py:168-191): numerical vs analytic gradient
    inp = (x.cuda().requires_grad_(), *cuda(tg, ll, tl))
    assert torch.autograd.gradcheck(crit, inp, eps=1e-6, atol=1e-4, nondet_tol=1e-6)


def test_bf16_golden(e2e):
    g = gold("bf16_c3_b4")
    xb = T(g["x_bf16_bits"]).view(torch.bfloat16)
    tg, ll, tl = cuda(T(g["targets"]), T(g["logits_lengths"]), T(g["targets_lengths"]))
    leaf = xb.cuda().requires_grad_()
    loss = e2e.CTCLoss(reduce=True, size_average=True)(leaf, tg, ll, tl)
    loss.backward()
    assert loss.dtype == torch.bfloat16 and leaf.grad.dtype == torch.bfloat16
    # bf16-stored outputs: one bf16 ulp of the reference evaluated on logits.float()
    assert_parity(loss, T(g["m0_loss"]), rtol=BF16_RTOL, atol=ATOL, what="bf16 loss")
    assert_parity(leaf.grad, T(g["m0_grad"]), rtol=BF16_RTOL, atol=ATOL, what="bf16 grad")
    # the arithmetic itself is fp32/fp64: same logits widened to fp32 meet the fp32 tolerance
    leaf32 = xb.float().cuda().requires_grad_()
    per_utt = e2e.CTCLoss()(leaf32, tg, ll, tl)
    assert_parity(per_utt, T(g["per_utt_loss"]), what="bf16->fp32 per-utterance loss")


# --------------------------------------------------------------------------------------------
# seeded synthetic draws of the BASELINE configs against the live oracle
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg,B,scale", [("c1", 4, 1.0), ("c2", 64, 1.0), ("c2", 16, 5.0), ("c3", 256, 1.0),
                                         ("c4", 32, 1.0), ("c5", 6, 1.0), ("c5", 3, 5.0)])
def test_config_vs_oracle(e2e, cfg, B, scale):
    _, T_, V, Lmin, Lmax, seed, dtype, full = oracle.CONFIGS[cfg]
    x, tg, ll, tl = oracle.make_inputs(B, T_, V, Lmin, Lmax, seed, dtype=dtype, full_length=full, scale=scale)
    ref_leaf = x.float().clone().requires_grad_()
    ref_loss = oracle.ctc_loss_module(oracle.engine(0), ref_leaf, tg, ll, tl, reduce=True, size_average=True)
    ref_loss.backward()
    leaf = x.cuda().requires_grad_()
    loss = e2e.CTCLoss(reduce=True, size_average=True, after_logsoftmax=False)(leaf, *cuda(tg, ll, tl))
    loss.backward()
    rtol = BF16_RTOL if dtype == torch.bfloat16 else RTOL
    assert_parity(loss, ref_loss, rtol=rtol, what=cfg + " loss")
    assert_parity(leaf.grad, ref_leaf.grad, rtol=rtol, what=cfg + " grad")
    # per-utterance losses and the engine contract (log-prob input, padding rows = exp(lp))
    lp = torch.log_softmax(x.float(), 2)
    l_ref, g_ref = oracle.engine(0).compute(lp, tg, ll, tl)
    l_gpu, g_gpu = e2e.CTCLossEngine(0).compute(lp.cuda(), *cuda(tg, ll, tl))
    assert_parity(l_gpu, l_ref, what=cfg + " engine losses")
    assert_parity(g_gpu, g_ref, what=cfg + " engine grads")


def _with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


# lattice kernels (end2end_b200/csrc): 0 the general kernel (every shape), 1 the wave kernel (latency shapes),
# 2 the one-warp-per-sweep kernel (throughput shapes); -1 the library's own dispatch
GENERAL, WAVE, SWEEP, AUTO = 0, 1, 2, -1


def _with_kernel(kind, fn):
    from end2end_b200 import _lib
    _lib.force_kernel(kind)
    try:
        return fn()
    finally:
        _lib.force_kernel(AUTO)


def _logits_grad_ref(g_ref, l_ref, ll):
    """What the reference's leaf gradient is for raw-logit input: softmax - posterior on valid frames, 0 on
    padding frames (log_softmax backward), all NaN for an infeasible utterance."""
    g_exp = g_ref.clone()
    for row, n in enumerate(ll.tolist()):
        g_exp[row, n:] = 0
        if not torch.isfinite(l_ref[row]):
            g_exp[row] = float("nan")
    return g_exp


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("kind", [GENERAL, WAVE, SWEEP])
def test_every_lattice_shape_vs_oracle(e2e, kind, fused):
    """Target lengths 0..70 (1..141 lattice cells) so every lane-edge placement of the entry and exit cells is
    hit, with repeats and short T; every lattice kernel, once through the fused single-kernel (dense) mode and
    once through row-stats + gather lattice + gradient kernels (the wave kernel has no gather mode: the library
    then dispatches on its own)."""
    g = torch.Generator().manual_seed(7)
    B, T_, V = 71, 90, 7
    x = torch.randn(B, T_, V, generator=g)
    tl = torch.arange(B)
    tg = torch.randint(1, V, (B, 70), generator=g)
    tg[::3, 1::2] = tg[::3, 0:-1:2]                       # plenty of adjacent repeats
    ll = torch.randint(T_ // 2, T_ + 1, (B,), generator=g)
    lp = torch.log_softmax(x, 2)
    l_ref, g_ref = oracle.engine(0).compute(lp, tg, ll, tl)
    for Lmax in (0, 1, 15, 31, 47, 63, 70):               # widths of the targets matrix: every cells-per-lane variant
        keep = tl <= Lmax
        for from_logits, inp in ((False, lp), (True, x)):
            args = (inp[keep].cuda(), *cuda(tg[keep][:, :Lmax], ll[keep], tl[keep]))

            def run():
                eng = e2e.CTCLossEngine(0)
                if fused:
                    return eng.compute(*args, from_logits=from_logits)
                losses, st = eng.forward(*args, from_logits=from_logits)
                return losses, eng.backward(st)
            l_gpu, g_gpu = _with_kernel(kind, run)
            what = "kernel %d fused %s Lmax %d" % (kind, fused, Lmax)
            assert_parity(l_gpu, l_ref[keep], what=what + " losses")
            g_exp = _logits_grad_ref(g_ref[keep], l_ref[keep], ll[keep]) if from_logits else g_ref[keep]
            assert_parity(g_gpu, g_exp, what=what + " grads (from_logits=%s)" % from_logits)


# (kernel, cells the variant covers): wave (K=4, NW=1/2/4) and the general kernel's classes (1, 2, 4, 6/8/10
# block rows per lane) and the sweep kernel's widest variants
LATTICE_VARIANTS = [(WAVE, 128), (WAVE, 256), (WAVE, 512), (GENERAL, 128), (GENERAL, 256), (GENERAL, 512),
                    (GENERAL, 768), (GENERAL, 1024), (GENERAL, 1280), (SWEEP, 512), (SWEEP, 1280)]


@pytest.mark.parametrize("kind,cells", LATTICE_VARIANTS)
def test_lattice_variant_every_shape_vs_oracle(e2e, kind, cells):
    """Target lengths spread over the variant's whole range (so the exit cells land on every lane / block edge,
    lanes beyond the lattice idle, the mirrored backward sweep pairs cells across blocks), adjacent repeats, ragged T
    incl. T_i = 1 and T_i == L + repeats (the only alignment), infeasible rows."""
    Lcap = (cells - 1) // 2
    g = torch.Generator().manual_seed(100 + cells)
    B, V = 24, 7
    tl = torch.linspace(0, Lcap, B).long()
    T_ = int(Lcap * 1.3) + 12
    tg = torch.randint(1, V, (B, max(Lcap, 1)), generator=g)
    tg[::3, 1::2] = tg[::3, 0:-1:2]                       # plenty of adjacent repeats
    rep = torch.tensor([int((tg[b, 1:tl[b]] == tg[b, :max(int(tl[b]) - 1, 0)]).sum()) for b in range(B)])
    ll = torch.randint(T_ // 2, T_ + 1, (B,), generator=g)
    ll = torch.minimum(torch.maximum(ll, tl + rep), torch.tensor(T_))
    ll[1] = 1                                             # a one-frame utterance with targets (infeasible)
    ll[5] = min(int(tl[5] + rep[5]), T_)                  # exactly one alignment
    ll[0] = 1                                             # L = 0, T = 1
    ll[7] = max(int(tl[7] + rep[7]) - 1, 1)               # infeasible by one frame
    x = torch.randn(B, T_, V, generator=g)
    lp = torch.log_softmax(x, 2)
    l_ref, g_ref = oracle.engine(0).compute(lp, tg, ll, tl)
    for from_logits, inp in ((False, lp), (True, x)):
        l_gpu, g_gpu = _with_kernel(kind, lambda: e2e.CTCLossEngine(0).compute(inp.cuda(), *cuda(tg, ll, tl), from_logits=from_logits))
        what = "kernel %d cells %d" % (kind, cells)
        assert_parity(l_gpu, l_ref, what=what + " losses")
        g_exp = _logits_grad_ref(g_ref, l_ref, ll) if from_logits else g_ref
        assert_parity(g_gpu, g_exp, what=what + " grads (from_logits=%s)" % from_logits)


@pytest.mark.parametrize("kind", [GENERAL, WAVE])
@pytest.mark.parametrize("scale", [1.0, 5.0, 12.0])
def test_lattice_peaky_and_dtypes(e2e, kind, scale):
    """Peaky emissions (the block exponents move tens of bits per frame and massless blocks must pick up the
    front's scale), 16-bit logits, time-major strides, blank != 0, int32 index tensors."""
    x, tg, ll, tl = oracle.make_inputs(6, 300, 29, 60, 140, 31, scale=scale)
    blank = 3
    tg = torch.where(tg == blank, torch.tensor(0), tg)
    l_ref, g_ref = oracle.engine(blank).compute(torch.log_softmax(x, 2), tg, ll, tl)
    x_tm = x.permute(1, 0, 2).contiguous().cuda()
    l_gpu, g_gpu = _with_kernel(kind, lambda: e2e.CTCLossEngine(blank).compute(
        torch.log_softmax(x_tm, 2).permute(1, 0, 2), tg.int().cuda(), ll.int().cuda(), tl.int().cuda()))
    assert_parity(l_gpu, l_ref, what="peaky x%g losses" % scale)
    assert_parity(g_gpu, g_ref, what="peaky x%g grads" % scale)
    for dt in (torch.bfloat16, torch.float16):
        xh = x.to(dt)
        lr, gr = oracle.engine(blank).compute(torch.log_softmax(xh.float(), 2), tg, ll, tl)
        lg, gg = _with_kernel(kind, lambda: e2e.CTCLossEngine(blank).compute(xh.cuda(), *cuda(tg, ll, tl), from_logits=True))
        for row, n in enumerate(ll.tolist()):
            gr[row, n:] = 0
        tol = BF16_RTOL if dt == torch.bfloat16 else 2.0 ** -11
        assert_parity(lg.float(), lr, rtol=tol, atol=tol, what="%s losses" % dt)
        assert_parity(gg.float(), gr, rtol=tol, atol=tol, what="%s grads" % dt)


def test_dispatch_is_bitwise_reproducible_and_kernels_agree(e2e):
    """A latency shape runs the wave kernel: the result is bitwise reproducible and equals the other two
    kernels' within the parity budget."""
    x, tg, ll, tl = oracle.make_inputs(16, 200, 29, 40, 100, 17)
    args = (x.cuda(), *cuda(tg, ll, tl))
    l_w, g_w = e2e.CTCLossEngine(0).compute(*args, from_logits=True)
    l_w2, g_w2 = e2e.CTCLossEngine(0).compute(*args, from_logits=True)
    assert torch.equal(l_w, l_w2) and torch.equal(g_w, g_w2)
    for kind in (GENERAL, SWEEP):
        l_k, g_k = _with_kernel(kind, lambda: e2e.CTCLossEngine(0).compute(*args, from_logits=True))
        l_k2, g_k2 = _with_kernel(kind, lambda: e2e.CTCLossEngine(0).compute(*args, from_logits=True))
        assert torch.equal(l_k, l_k2) and torch.equal(g_k, g_k2)          # integer accumulation: no run-to-run noise
        assert_parity(l_w, l_k, what="dispatch vs kernel %d losses" % kind)
        assert_parity(g_w, g_k, what="dispatch vs kernel %d grads" % kind)


@pytest.mark.parametrize("kind", [GENERAL, WAVE])
def test_tight_peaky_alignments(e2e, kind):
    """T_i == L_i + repeats (exactly one alignment) under very peaky emissions (logits x10): the single
    feasible path runs along the mass front, tens of orders of magnitude below the dead-end mass behind it.
    The per-block exponents must keep it (found by the differential fuzz in round 1)."""
    g = torch.Generator().manual_seed(5)
    B, V = 6, 5
    tl = torch.tensor([55, 73, 66, 40, 120, 9])
    T_ = 170
    tg = torch.randint(1, V, (B, 120), generator=g)
    ll = torch.zeros(B, dtype=torch.int64)
    for b in range(B):
        L = int(tl[b])
        rep = int((tg[b, 1:L] == tg[b, :L - 1]).sum())
        ll[b] = L + rep + (b % 3)                      # tight, or one/two spare frames
    assert int(ll.max()) <= T_
    x = torch.randn(B, T_, V, generator=g) * 10.0
    lp = torch.log_softmax(x, 2)
    l_ref, g_ref = oracle.engine(0).compute(lp, tg, ll, tl)
    assert torch.isfinite(l_ref).all()
    l_gpu, g_gpu = _with_kernel(kind, lambda: e2e.CTCLossEngine(0).compute(lp.cuda(), *cuda(tg, ll, tl)))
    assert_parity(l_gpu, l_ref, what="tight peaky losses")
    assert_parity(g_gpu, g_ref, what="tight peaky grads")


def test_single_label_symbol_tight_peaky(e2e):
    """V = 2 (one label, so every target is a repeat) in exactly 2L-1 / 2L+1 frames under logits x10: round 1's
    open corner of the one-warp-per-sweep kernel (its lane-exponent rule lost the only path: gradient off by 1.0).
    The library now runs such alphabets on the general kernel, for every batch size; arbitrated by the oracle."""
    for B, T_ in ((4, 97), (200, 61)):                   # a latency shape and a throughput shape
        g = torch.Generator().manual_seed(1000 + B)
        L = (T_ - 1) // 2
        x = torch.randn(B, T_, 2, generator=g) * 10.0
        tg = torch.ones(B, L + 1, dtype=torch.int64)
        tl = torch.full((B,), L, dtype=torch.int64)
        tl[1::2] = L + 1                                  # 2L+1 frames for L labels, 2L'-1 for L' = L+1
        ll = torch.full((B,), T_, dtype=torch.int64)
        for from_logits in (True, False):
            inp = x if from_logits else torch.log_softmax(x, 2)
            l_ref, g_ref = oracle.engine(0).compute(torch.log_softmax(x, 2), tg, ll, tl)
            assert torch.isfinite(l_ref).all()
            l_gpu, g_gpu = e2e.CTCLossEngine(0).compute(inp.cuda(), *cuda(tg, ll, tl), from_logits=from_logits)
            assert_parity(l_gpu, l_ref, what="V=2 B=%d losses" % B)
            assert_parity(g_gpu, _logits_grad_ref(g_ref, l_ref, ll) if from_logits else g_ref, what="V=2 B=%d grads" % B)


@pytest.mark.parametrize("kind", [GENERAL, WAVE, SWEEP])
def test_minus_infinity_inputs(e2e, kind):
    """A masked symbol (-inf logit / log-prob): an exact zero emission, as in the reference's log_sum_exp
    (math_utils.h:8-16), not a NaN -- both as raw logits and as log-probabilities, also when the masked
    symbol is a label (infeasible: +inf loss, NaN block)."""
    x = torch.randn(4, 20, 6, generator=torch.Generator().manual_seed(9))
    x[:, :, 4] = float("-inf")
    tg = torch.tensor([[1, 2, 3], [1, 1, 2], [4, 2, 0], [5, 3, 1]])
    tl = torch.tensor([3, 3, 2, 3])
    ll = torch.tensor([20, 15, 9, 20])
    for from_logits in (True, False):
        inp = x if from_logits else torch.log_softmax(x, 2)
        l_ref, g_ref = oracle.engine(0).compute(torch.log_softmax(x, 2), tg, ll, tl)
        assert torch.isposinf(l_ref[2]) and torch.isfinite(l_ref[[0, 1, 3]]).all()
        l_gpu, g_gpu = _with_kernel(kind, lambda: e2e.CTCLossEngine(0).compute(inp.cuda(), *cuda(tg, ll, tl), from_logits=from_logits))
        assert_parity(l_gpu, l_ref, what="-inf losses (from_logits=%s)" % from_logits)
        assert_parity(g_gpu, _logits_grad_ref(g_ref, l_ref, ll) if from_logits else g_ref,
                      what="-inf grads (from_logits=%s)" % from_logits)


def test_longest_supported_targets_and_limit(e2e):
    """639 labels (1279 lattice cells) is the build limit: the general kernel's widest class and the sweep
    kernel's widest variant; one more label is rejected with NotImplementedError, not computed wrongly."""
    g = torch.Generator().manual_seed(11)
    B, T_, V = 3, 700, 6
    x = torch.randn(B, T_, V, generator=g)
    tl = torch.tensor([639, 513, 600])
    tg = 1 + (torch.arange(639)[None, :] * 2 + torch.arange(B)[:, None]) % (V - 1)        # no adjacent repeats: all feasible
    ll = torch.tensor([700, 690, 650])
    lp = torch.log_softmax(x, 2)
    l_ref, g_ref = oracle.engine(0).compute(lp, tg, ll, tl)
    assert torch.isfinite(l_ref).all()
    for kind in (GENERAL, SWEEP):
        l_gpu, g_gpu = _with_kernel(kind, lambda: e2e.CTCLossEngine(0).compute(lp.cuda(), *cuda(tg, ll, tl)))
        assert_parity(l_gpu, l_ref, what="L=639 losses kernel %d" % kind)
        assert_parity(g_gpu, g_ref, what="L=639 grads kernel %d" % kind)
    with pytest.raises(NotImplementedError):
        e2e.CTCLossEngine(0).compute(lp.cuda(), torch.ones(B, 640, dtype=torch.int64).cuda(), ll.cuda(), tl.cuda())


def test_split_forward_backward_and_repeated_backward(e2e):
    """engine.forward()/backward() (row stats + lattice, then the gradient kernel with grad_output
    folded in) against engine.step(); and two backward passes through a retained graph."""
    x, tg, ll, tl = oracle.make_inputs(16, 120, 29, 20, 50, 5)
    eng = e2e.CTCLossEngine(0)
    xc, tgc, llc, tlc = cuda(x, tg, ll, tl)
    losses, state = eng.forward(xc, tgc, llc, tlc, from_logits=True)
    go = torch.rand(16, device="cuda") + 0.5
    g_split = eng.backward(state, go, 0.25)
    l_step, g_step, red, pair = eng.step(xc, tgc, llc, tlc, from_logits=True, grad_scale=0.25, reduce_scale=1.0 / 16,
                                         want_pair=True)
    eng.scale_rows_(g_step, go)
    assert_parity(l_step, losses, what="step vs split losses")
    assert_parity(g_step, g_split, what="step vs split grads")
    assert abs(red.item() - losses.double().mean().item()) < 1e-3 and pair[1].item() == 16.0
    assert abs(pair[0].item() - losses.double().mean().item()) < 1e-3    # the pair carries the scaled sum
    leaf = xc.clone().requires_grad_()
    loss = e2e.CTCLoss(reduce=True, size_average=False)(leaf, tgc, llc, tlc)
    loss.backward(torch.tensor(2.0, device="cuda"), retain_graph=True)
    g1 = leaf.grad.clone()
    leaf.grad = None
    loss.backward(torch.tensor(3.0, device="cuda"))
    assert_parity(leaf.grad * 2.0, g1 * 3.0, what="repeated backward")


def test_time_major_in_place_and_reduce_modes(e2e):
    x, tg, ll, tl = oracle.make_inputs(6, 40, 12, 3, 9, 5)
    x_tm = x.permute(1, 0, 2).contiguous()
    for kw in (dict(), dict(reduce=True), dict(reduce=True, size_average=True), dict(size_average=True)):
        ref_leaf = x_tm.clone().requires_grad_()
        ref = oracle.ctc_loss_module(oracle.engine(0), ref_leaf, tg, ll, tl, time_major=True, **kw)
        (ref.sum() if ref.dim() else ref).backward()
        leaf = x_tm.cuda().requires_grad_()
        out = e2e.CTCLoss(time_major=True, **kw)(leaf, *cuda(tg, ll, tl))
        (out.sum() if out.dim() else out).backward()
        assert out.shape == ref.shape                                    # [B] or 0-dim, never (1,)
        assert_parity(out, ref, what="tm loss %s" % kw)
        assert_parity(leaf.grad, ref_leaf.grad, what="tm grad %s" % kw)
        assert leaf.grad.shape == x_tm.shape and leaf.grad.is_contiguous()


def test_upstream_gradient_scaling_and_nan_through_zero(e2e):
    x, tg, ll, tl = oracle.make_inputs(5, 12, 6, 2, 5, 9)
    tg[1, :3] = 2
    tl[1], ll[1] = 3, 3                                                   # infeasible: needs 5 frames
    w = torch.tensor([0.5, 0.0, -2.0, 3.0, 1.0])
    ref_leaf = x.clone().requires_grad_()
    (oracle.ctc_loss_module(oracle.engine(0), ref_leaf, tg, ll, tl)[[0, 2, 3, 4]] * w[[0, 2, 3, 4]]).sum().backward()
    leaf = x.cuda().requires_grad_()
    loss = e2e.CTCLoss()(leaf, *cuda(tg, ll, tl))
    assert torch.isposinf(loss[1])
    loss.backward(w.cuda())                                               # zero upstream grad on the inf utterance
    assert torch.isnan(leaf.grad[1]).all()                                # NaN survives * 0, as in the reference
    keep = [0, 2, 3, 4]
    assert_parity(leaf.grad[keep], ref_leaf.grad[keep], what="weighted grad")


def test_non_contiguous_views_and_int32(e2e):
    x, tg, ll, tl = oracle.make_inputs(4, 30, 10, 2, 8, 13)
    big = torch.zeros(4, 30, 16)
    big[:, :, 3:13] = x
    view = big.cuda()[:, :, 3:13]                                        # unit alphabet stride, padded rows
    lp_ref = torch.log_softmax(x, 2)
    l_ref, g_ref = oracle.engine(0).compute(lp_ref, tg, ll, tl)
    lp_view = torch.log_softmax(view, 2)
    big_lp = torch.zeros(4, 30, 16, device="cuda")
    big_lp[:, :, 3:13] = lp_view
    l, g_ = e2e.CTCLossEngine(0).compute(big_lp[:, :, 3:13], tg.int().cuda(), ll.int().cuda(), tl.int().cuda())
    assert_parity(l, l_ref, what="view losses")
    assert_parity(g_, g_ref, what="view grads")
    l2, g2 = e2e.CTCLossEngine(0).compute(lp_ref.cuda().transpose(1, 2).contiguous().transpose(1, 2),
                                          *cuda(tg, ll, tl))              # alphabet stride != 1 -> repacked
    assert_parity(g2, g_ref, what="repacked grads")


def test_invalid_arguments_rejected(e2e):
    x, tg, ll, tl = oracle.make_inputs(3, 10, 5, 1, 3, 1)
    crit = e2e.CTCLoss()
    with pytest.raises(ValueError):
        crit(x.cuda(), tg, torch.tensor([10, 0, 10]), tl)               # T_i = 0
    with pytest.raises(ValueError):
        crit(x.cuda(), tg, torch.tensor([10, 11, 10]), tl)              # T_i > T
    with pytest.raises(ValueError):
        crit(x.cuda(), tg, ll, torch.tensor([1, 4, 1]))                 # L_i > Lmax
    with pytest.raises(ValueError):
        crit(x.cuda(), torch.full_like(tg, 5), ll, tl)                  # label >= V
    with pytest.raises(ValueError):
        e2e.CTCLoss(blank_idx=5)(x.cuda(), tg, ll, tl)                  # blank outside the alphabet
    # device-resident lengths are checked on the device: flagged, NaN results, no crash
    eng = e2e.CTCLossEngine(0)
    losses, state = eng.forward(x.cuda(), tg.cuda(), torch.tensor([10, 0, 10]).cuda(), tl.cuda(), True)
    assert eng.check(state) & 1
    assert torch.isnan(losses[1]) and torch.isfinite(losses[[0, 2]]).all()


# --------------------------------------------------------------------------------------------
# greedy decoder: bit-exact
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["simple", "sm", "probs_1", "probs_2"])
def test_greedy_known_answers(e2e, name):
    g = gold("kat_greedy")
    ll = g[name + "_ll"]
    lengths = None if ll[0] < 0 else T(ll)
    dec = e2e.CTCDecoder(beam_width=1, blank_idx=int(g[name + "_blank"]), labels=[str(s) for s in g[name + "_labels"]],
                         time_major=False)
    for dev in ("cpu", "cuda"):
        r = dec.decode(T(g[name + "_x"]).to(dev), lengths)
        assert r.decoded_sentences == [str(s) for s in g[name + "_sentences"]]
        assert r.decoded_targets.dtype == torch.int64 and not r.decoded_targets.is_cuda
        assert torch.equal(r.decoded_targets, T(g[name + "_targets"]))
        assert torch.equal(r.decoded_targets_lengths, T(g[name + "_lengths"]))
        r2 = dec.decode_greedy(T(g[name + "_x"]).to(dev), lengths)
        assert torch.equal(r2.decoded_targets, r.decoded_targets)


def test_greedy_ties_nan_bf16_time_major(e2e):
    g = gold("greedy_random")
    x, ll, blank = T(g["x"]), T(g["ll"]), int(g["blank"])
    dec = e2e.CTCDecoder(beam_width=1, blank_idx=blank)
    for dev in ("cpu", "cuda"):
        r = dec.decode(x.to(dev), ll.to(dev))
        assert torch.equal(r.decoded_targets, T(g["targets"])) and torch.equal(r.decoded_targets_lengths, T(g["lengths"]))
        r = dec.decode(x.to(dev))                                         # lengths None -> all frames
        assert torch.equal(r.decoded_targets, T(g["targets_full"])) and torch.equal(r.decoded_targets_lengths, T(g["lengths_full"]))
        assert r.decoded_sentences == [""] * x.size(0)
    xb = T(g["xb_bits"]).view(torch.bfloat16)
    for dev in ("cpu", "cuda"):
        r = e2e.CTCDecoder(beam_width=1).decode(xb.to(dev), T(g["llb"]))
        assert torch.equal(r.decoded_targets, T(g["targets_b"])) and torch.equal(r.decoded_targets_lengths, T(g["lengths_b"]))
        r = e2e.CTCDecoder(beam_width=1, time_major=True).decode(xb.transpose(0, 1).contiguous().to(dev), T(g["llb"]))
        assert torch.equal(r.decoded_targets, T(g["targets_b"]))
        r = e2e.CTCDecoder(beam_width=1, time_major=True).decode(xb.to(dev).transpose(0, 1), T(g["llb"]))  # strided view
        assert torch.equal(r.decoded_targets, T(g["targets_b"]))


def test_greedy_c4_full_size_vs_oracle(e2e):
    B, T_, V, Lmin, Lmax, seed, _, _ = oracle.CONFIGS["c4"]
    x, _, ll, _ = oracle.make_inputs(B, T_, V, Lmin, Lmax, seed)
    ref = oracle.greedy_decode(x, ll)
    r = e2e.CTCDecoder(beam_width=1).decode(x.cuda(), ll.cuda())
    assert torch.equal(r.decoded_targets, ref[0]) and torch.equal(r.decoded_targets_lengths, ref[1])
    r16 = e2e.CTCDecoder(beam_width=1).decode(x.half().cuda(), ll)
    ref16 = oracle.greedy_decode(x.half(), ll)
    assert torch.equal(r16.decoded_targets, ref16[0])
    with pytest.raises(NotImplementedError):      # KenLM decoding stays the reference's CPU code
        e2e.CTCDecoder(beam_width=20, lm_path="/no/such/model.arpa").decode(x.cuda())


# --------------------------------------------------------------------------------------------
# BASELINE full sizes: size-independent properties
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg,B", [("c3", 1024), ("c4", 128), ("c5", 96)])
def test_full_size_properties(e2e, cfg, B):
    _, T_, V, Lmin, Lmax, seed, dtype, full = oracle.CONFIGS[cfg]
    x, tg, ll, tl = oracle.make_inputs(B, T_, V, Lmin, Lmax, seed, dtype=torch.float32, full_length=full)
    leaf = x.cuda().requires_grad_()
    tgc, llc, tlc = cuda(tg, ll, tl)
    per_utt = e2e.CTCLoss()(leaf, tgc, llc, tlc)
    per_utt.sum().backward()
    grad = leaf.grad
    assert torch.isfinite(per_utt).all() and torch.isfinite(grad).all()
    valid = (torch.arange(T_, device="cuda")[None, :] < llc[:, None])
    # (1) softmax and posterior both sum to one over the alphabet: gradient rows sum to ~0
    assert float(grad.sum(2).abs().max()) < 2e-4
    # (2) padding frames get exactly zero (fused log-softmax backward), valid frames do not
    assert float(grad[~valid].abs().max()) == 0.0
    # (3) blank posterior mass: d loss / d logit[blank] = softmax - posterior, bounded by 1 in magnitude
    assert float(grad.abs().max()) <= 1.0 + 1e-5
    # (4) the reduced losses are the sum / mean of the per-utterance losses
    s = e2e.CTCLoss(reduce=True)(x.cuda(), tgc, llc, tlc)
    m = e2e.CTCLoss(reduce=True, size_average=True)(x.cuda(), tgc, llc, tlc)
    assert abs(s.item() - per_utt.double().sum().item()) <= 1e-6 * abs(s.item())
    assert abs(m.item() - per_utt.double().mean().item()) <= 1e-6 * abs(m.item())
    # (5) the loss of an utterance does not depend on its batch neighbours or on padding frames
    idx = torch.tensor([0, B // 2, B - 1])
    sub = e2e.CTCLoss()(x[idx].cuda(), tg[idx].cuda(), ll[idx].cuda(), tl[idx].cuda())
    assert torch.equal(sub, per_utt[idx.cuda()])
    # (6) spot check against the oracle
    lp = torch.log_softmax(x[idx], 2)
    l_ref, _ = oracle.engine(0).compute(lp, tg[idx], ll[idx], tl[idx])
    assert_parity(sub, l_ref, what=cfg + " spot losses")


def _oracle_sub_batched(x, tg, ll, tl, sub):
    """The oracle engine over a large batch, `sub` utterances at a time (the reference keeps ~46 MB of fp64 lattices
    live per long-form utterance: SURVEY.md 8a, a12)."""
    losses, grads = [], []
    for i in range(0, x.size(0), sub):
        sl = slice(i, i + sub)
        l_, g_ = oracle.engine(0).compute(torch.log_softmax(x[sl].float(), 2), tg[sl], ll[sl], tl[sl])
        losses.append(l_); grads.append(g_)
    return torch.cat(losses), torch.cat(grads)


@pytest.mark.parametrize("cfg,B,sub", [("c3", 1024, 256), ("c4", 128, 32), ("c5", 256, 32)])
def test_full_size_gradient_parity(e2e, cfg, B, sub):
    """Loss AND gradient of every utterance at BASELINE's full batch sizes (c5: 256 of its 2048 utterances, the
    oracle sub-batched) against the compiled reference, at the parity tolerance."""
    _, T_, V, Lmin, Lmax, seed, dtype, full = oracle.CONFIGS[cfg]
    x, tg, ll, tl = oracle.make_inputs(B, T_, V, Lmin, Lmax, seed, dtype=dtype, full_length=full)
    l_ref, g_ref = _oracle_sub_batched(x, tg, ll, tl, sub)
    g_ref = _logits_grad_ref(g_ref, l_ref, ll)
    l_gpu, g_gpu = e2e.CTCLossEngine(0).compute(x.cuda(), *cuda(tg, ll, tl), from_logits=True)
    if dtype == torch.bfloat16:   # bf16-stored results within one bf16 ulp of the reference run on logits.float()
        assert_parity(l_gpu.float(), l_ref, rtol=BF16_RTOL, what=cfg + " losses")
        assert_parity(g_gpu.float(), g_ref, rtol=BF16_RTOL, what=cfg + " grads")
    else:
        assert_parity(l_gpu, l_ref, what=cfg + " losses")
        assert_parity(g_gpu, g_ref, what=cfg + " grads")


def test_c5_full_batch_runs_and_matches_on_a_sample(e2e):
    """BASELINE config 5 at its full size (B=2048, T=1600, L<=600: a ~25 GB workspace): finite losses, zero padding
    rows, gradient rows summing to zero, and loss + gradient parity against the oracle on 48 utterances spread over
    the batch (an utterance's result does not depend on its batch neighbours)."""
    B, T_, V, Lmin, Lmax, seed, dtype, full = oracle.CONFIGS["c5"]
    x, tg, ll, tl = oracle.make_inputs(B, T_, V, Lmin, Lmax, seed, dtype=dtype, full_length=full)
    xc = x.cuda()
    l_gpu, g_gpu = e2e.CTCLossEngine(0).compute(xc, *cuda(tg, ll, tl), from_logits=True)
    assert torch.isfinite(l_gpu).all() and torch.isfinite(g_gpu).all()
    valid = torch.arange(T_, device="cuda")[None, :] < ll.cuda()[:, None]
    assert float(g_gpu[~valid].abs().max()) == 0.0
    assert float(g_gpu.sum(2).abs().max()) < 2e-4
    idx = torch.arange(0, B, B // 48)[:48]
    l_ref, g_ref = _oracle_sub_batched(x[idx], tg[idx], ll[idx], tl[idx], 16)
    assert_parity(l_gpu[idx.cuda()], l_ref, what="c5 B=2048 sample losses")
    assert_parity(g_gpu[idx.cuda()], _logits_grad_ref(g_ref, l_ref, ll[idx]), what="c5 B=2048 sample grads")
    del g_gpu, xc
    torch.cuda.empty_cache()


def test_greedy_idempotence_full_size(e2e):
    """Decoding the one-hot re-encoding of a collapsed greedy path returns the same labels."""
    B, T_, V, *_ = oracle.CONFIGS["c3"]
    x = torch.randn(B, T_, V, generator=torch.Generator().manual_seed(2)).cuda()
    dec = e2e.CTCDecoder(beam_width=1)
    r = dec.decode(x)
    onehot = torch.zeros(B, T_, V, device="cuda")
    onehot[:, :, 0] = 0.5                                                 # blank wins wherever nothing is set
    tgt = r.decoded_targets.cuda()
    pos = torch.arange(T_, device="cuda")[None, :].expand(B, T_)
    mask = pos < r.decoded_targets_lengths.cuda()[:, None]
    # place label i at frame 2i (blank between) when it fits, else skip the check for that row
    fits = r.decoded_targets_lengths * 2 <= T_
    rows = torch.nonzero(fits.cuda()).flatten()
    for b in rows[:64].tolist():
        n = int(r.decoded_targets_lengths[b])
        onehot[b, torch.arange(n, device="cuda") * 2, tgt[b, :n]] = 1.0
    r2 = dec.decode(onehot)
    for b in rows[:64].tolist():
        n = int(r.decoded_targets_lengths[b])
        assert int(r2.decoded_targets_lengths[b]) == n
        assert torch.equal(r2.decoded_targets[b, :n], r.decoded_targets[b, :n])
    assert bool(mask.any())


# --------------------------------------------------------------------------------------------
# the raw C ABI (what a non-Python host binds)
# --------------------------------------------------------------------------------------------
def test_c_abi_direct_calls(e2e):
    from end2end_b200 import _lib
    L = _lib.load()
    x, tg, ll, tl = oracle.make_inputs(4, 50, 28, 10, 29, 0, full_length=True)
    lp = torch.log_softmax(x, 2).cuda()
    tgc, llc, tlc = cuda(tg, ll, tl)
    d = _lib.Desc()
    d.batch, d.max_frames, d.alphabet, d.max_targets = 4, 50, 28, 29
    d.blank_idx, d.dtype, d.targets_itype, d.lengths_itype, d.from_logits = 0, _lib.E2E_F32, _lib.E2E_I64, _lib.E2E_I64, 0
    d.logits_stride_b, d.logits_stride_t = 50 * 28, 28
    d.grads_stride_b, d.grads_stride_t = 50 * 28, 28
    d.targets_stride_b = 29
    n = L.e2e_ctc_loss_workspace_bytes(ctypes.byref(d))
    assert n > 0
    ws = torch.empty(n, dtype=torch.uint8, device="cuda")
    losses = torch.empty(4, device="cuda")
    grads = torch.empty_like(lp)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    before = _lib.launch_count()
    rc = L.e2e_ctc_loss_fwd_bwd_device(ctypes.byref(d), p(lp), p(tgc), p(llc), p(tlc), p(losses), p(grads), p(ws), n, stream)
    assert rc == 0, L.e2e_last_error_string()
    assert _lib.launch_count() - before == 1                                # V <= 128: one fused kernel
    l_ref, g_ref = oracle.engine(0).compute(lp.cpu(), tg, ll, tl)
    assert_parity(losses, l_ref, what="abi losses")
    assert_parity(grads, g_ref, what="abi grads")
    total = torch.empty((), device="cuda")
    pair = torch.empty(2, dtype=torch.float64, device="cuda")
    assert L.e2e_ctc_loss_reduce_device(p(losses), _lib.E2E_F32, 4, 0.25, p(total), p(pair), stream) == 0
    assert abs(total.item() - l_ref.double().mean().item()) < 1e-4 and pair[1].item() == 4.0
    # error paths: small / misaligned workspace, null pointers, bad descriptor
    assert L.e2e_ctc_loss_fwd_bwd_device(ctypes.byref(d), p(lp), p(tgc), p(llc), p(tlc), p(losses), p(grads), p(ws), 4096, stream) == 3
    assert L.e2e_ctc_loss_fwd_bwd_device(ctypes.byref(d), p(lp), p(tgc), p(llc), p(tlc), p(losses), p(grads),
                                         ctypes.c_void_p(ws.data_ptr() + 8), n, stream) == 3
    assert L.e2e_ctc_loss_fwd_bwd_device(ctypes.byref(d), None, p(tgc), p(llc), p(tlc), p(losses), p(grads), p(ws), n, stream) == 1
    assert b"null" in L.e2e_last_error_string()
    d.blank_idx = 28
    assert L.e2e_ctc_loss_workspace_bytes(ctypes.byref(d)) == 0
    torch.cuda.synchronize()


@pytest.mark.parametrize("cfg,B,time_major,reduce,mean", [("c1", 4, False, True, True), ("c2", 16, True, True, False),
                                                          ("c4", 6, False, False, False), ("c3", 64, False, True, True)])
def test_graphed_step_matches_autograd_and_oracle(e2e, cfg, B, time_major, reduce, mean):
    """criterion.graphed(...) captures forward + backward once (e2e_ctc_graph_*); every replay must
    give what the autograd path gives -- bitwise, it runs the same kernels -- and the oracle's result, also after
    the CONTENTS of the bound buffers changed (a training loop copies each batch into them)."""
    from end2end_b200 import _lib
    _, T_, V, Lmin, Lmax, seed, dtype, full = oracle.CONFIGS[cfg]
    crit = e2e.CTCLoss(reduce=reduce, size_average=mean, time_major=time_major)
    x, tg, ll, tl = oracle.make_inputs(B, T_, V, Lmin, Lmax, seed, dtype=dtype, full_length=full)
    buf = (x.permute(1, 0, 2).contiguous() if time_major else x).cuda()
    tgc, llc, tlc = cuda(tg, ll, tl)
    step = crit.graphed(buf, tgc, llc, tlc)
    for rnd in range(3):
        if rnd:      # a new batch into the SAME buffers
            x, tg, ll, tl = oracle.make_inputs(B, T_, V, Lmin, Lmax, seed + rnd, dtype=dtype, full_length=full)
            buf.copy_(x.permute(1, 0, 2) if time_major else x)
            tgc.copy_(tg); llc.copy_(ll); tlc.copy_(tl)
        before = _lib.launch_count()
        loss, grad = step.replay()
        assert _lib.launch_count() - before >= 1       # the graph's kernels are counted at launch
        leaf = buf.detach().clone().requires_grad_()
        ref = crit(leaf, tgc, llc, tlc)
        ref.sum().backward()
        assert torch.equal(loss, ref.detach()) and torch.equal(grad, leaf.grad), (cfg, rnd)
        xo = x.float().clone().requires_grad_()
        lo = oracle.ctc_loss_module(oracle.engine(0), xo, tg, ll, tl, reduce=reduce, size_average=mean)
        lo.sum().backward()
        go = xo.grad.permute(1, 0, 2) if time_major else xo.grad
        if dtype == torch.bfloat16:
            assert_parity(loss.float(), lo.detach(), rtol=BF16_RTOL, what=cfg + " graphed loss")
            assert_parity(grad.float(), go, rtol=BF16_RTOL, what=cfg + " graphed grad")
        else:
            assert_parity(loss, lo.detach(), what=cfg + " graphed loss")
            assert_parity(grad, go, what=cfg + " graphed grad")
    with pytest.raises(ValueError):
        crit.graphed(buf, tg, llc, tlc)                # host-resident targets cannot be bound into a graph


def test_sharded_graphed_step_single_rank(e2e):
    """ShardedCTCLoss.graphed on one rank (no process group): two alternating captured steps, same results as the
    module; the collective path is covered by tests/dist_check.py under torchrun."""
    from end2end_b200.distributed import ShardedCTCLoss
    x, tg, ll, tl = oracle.make_inputs(12, 90, 29, 5, 30, 5)
    xc, tgc, llc, tlc = cuda(x, tg, ll, tl)
    crit = ShardedCTCLoss(reduce=True, size_average=True, global_batch=24)
    step = crit.graphed(xc, tgc, llc, tlc)
    leaf = xc.clone().requires_grad_()
    ref = crit(leaf, tgc, llc, tlc)
    ref.backward()
    for _ in range(3):
        total, grad = step.replay()
        step.wait()
        assert torch.equal(total, ref.detach()) and torch.equal(grad, leaf.grad)
    with pytest.raises(ValueError):
        ShardedCTCLoss(reduce=True, size_average=True).graphed(xc, tgc, llc, tlc)     # a mean needs global_batch


def test_host_engine_traffic_and_pinned_buffers(e2e):
    x, tg, ll, tl = oracle.make_inputs(8, 60, 29, 5, 20, 17)
    eng = e2e.CTCLossEngine(0)
    lp = torch.log_softmax(x, 2).pin_memory()
    losses, grads = eng.compute(lp, tg, ll, tl)
    h2d, d2h = eng.last_host_traffic()
    assert h2d == lp.numel() * 4 + tg.numel() * 8 + 2 * 8 * 8 and d2h == lp.numel() * 4 + 8 * 4
    l_ref, g_ref = oracle.engine(0).compute(lp, tg, ll, tl)
    assert_parity(losses, l_ref, what="host losses")
    assert_parity(grads, g_ref, what="host grads")
    # the module on CPU tensors takes the same route (forward) and scales on the host (backward)
    leaf = x.clone().requires_grad_()
    loss = e2e.CTCLoss(reduce=True, size_average=True)(leaf, tg, ll, tl)
    loss.backward()
    ref_leaf = x.clone().requires_grad_()
    oracle.ctc_loss_module(oracle.engine(0), ref_leaf, tg, ll, tl, reduce=True, size_average=True).backward()
    assert_parity(leaf.grad, ref_leaf.grad, what="host module grad")


@pytest.mark.parametrize("chunks", [2, 3, 8])
def test_host_engine_chunked_pipeline(e2e, chunks):
    """The host entry point cuts a batch-major batch into chunks of utterances that flow through copy-in /
    kernels / copy-out on separate streams: same results as the single-stream path, ragged last chunk,
    infeasible rows, bf16; time-major host tensors take the single-stream path; bad lengths are rejected
    on the host before anything is enqueued."""
    x, tg, ll, tl = oracle.make_inputs(7, 60, 11, 3, 12, 77)
    ll[2] = 5                                                         # infeasible: T < L
    tl[2] = 9
    eng = e2e.CTCLossEngine(0)
    l_one, g_one = _with_env({"E2E_CTC_HOST_CHUNKS": "1"}, lambda: eng.compute(x, tg, ll, tl, from_logits=True))
    l_ch, g_ch = _with_env({"E2E_CTC_HOST_CHUNKS": str(chunks)}, lambda: eng.compute(x, tg, ll, tl, from_logits=True))
    assert not l_ch.is_cuda and torch.isinf(l_ch[2]) and torch.isnan(g_ch[2]).all()
    assert torch.equal(torch.nan_to_num(l_one, 0, 0, 0), torch.nan_to_num(l_ch, 0, 0, 0))
    assert torch.equal(torch.isnan(g_one), torch.isnan(g_ch))
    assert_parity(g_ch, g_one, what="chunked vs single-stream grads")
    l_ref, g_ref = oracle.engine(0).compute(torch.log_softmax(x, 2), tg, ll, tl)
    assert_parity(l_ch, l_ref, what="chunked host losses")
    xb = x.to(torch.bfloat16)
    lb1, gb1 = _with_env({"E2E_CTC_HOST_CHUNKS": "1"}, lambda: eng.compute(xb, tg, ll, tl, from_logits=True))
    lbc, gbc = _with_env({"E2E_CTC_HOST_CHUNKS": str(chunks)}, lambda: eng.compute(xb, tg, ll, tl, from_logits=True))
    assert torch.equal(torch.nan_to_num(gb1.float(), 0, 0, 0), torch.nan_to_num(gbc.float(), 0, 0, 0))
    x_tm = x.permute(1, 0, 2).contiguous().permute(1, 0, 2)           # time-major storage, batch-major view
    l_tm, g_tm = _with_env({"E2E_CTC_HOST_CHUNKS": str(chunks)}, lambda: eng.compute(x_tm, tg, ll, tl, from_logits=True))
    assert_parity(g_tm, g_one, what="time-major host grads")
    with pytest.raises(ValueError):
        _with_env({"E2E_CTC_HOST_CHUNKS": str(chunks)},
                  lambda: eng.compute(x, tg, torch.tensor([60, 61, 60, 60, 60, 60, 60]), tl, from_logits=True))
    with pytest.raises(ValueError):
        eng.compute(x, torch.full_like(tg, 11), ll, tl, from_logits=True)


@pytest.mark.parametrize("cfg,B", [("c1", 4), ("c2", 12), ("c4", 5), ("c3", 40)])
def test_host_engine_zero_copy_results_equal_staged(e2e, cfg, B):
    """Pinned result buffers are written by the kernels directly (no device->host copy stage); pageable or
    E2E_CTC_HOST_ZERO_COPY=0 takes the staged copies.  Same bits either way, single-stream and chunked."""
    _, T_, V, Lmin, Lmax, seed, dtype, full = oracle.CONFIGS[cfg]
    x, tg, ll, tl = oracle.make_inputs(B, T_, V, Lmin, Lmax, seed, dtype=dtype, full_length=full)
    eng = e2e.CTCLossEngine(0)
    pinned = [t.pin_memory() for t in (x, tg, ll, tl)]
    outs = []
    for env in ({"E2E_CTC_HOST_ZERO_COPY": "0", "E2E_CTC_HOST_CHUNKS": "1"}, {"E2E_CTC_HOST_CHUNKS": "1"},
                {"E2E_CTC_HOST_CHUNKS": "3"}, {"E2E_CTC_HOST_ZERO_COPY": "0", "E2E_CTC_HOST_CHUNKS": "3"}, {}):
        outs.append(_with_env(env, lambda: eng.compute(*pinned, from_logits=True)))
        outs.append(_with_env(env, lambda: eng.compute(pinned[0], tg, ll, tl, from_logits=True)))   # pageable index tensors
    for l_, g_ in outs[1:]:
        assert torch.equal(l_, outs[0][0]) and torch.equal(g_, outs[0][1])
    h2d, d2h = eng.last_host_traffic()
    assert d2h == x.numel() * x.element_size() + B * x.element_size()       # result bytes are counted either way
    l_ref, g_ref = _oracle_sub_batched(x, tg, ll, tl, 64)
    g_ref = _logits_grad_ref(g_ref, l_ref, ll)
    if dtype == torch.bfloat16:
        assert_parity(outs[0][1].float(), g_ref, rtol=BF16_RTOL, what=cfg + " host grads")
    else:
        assert_parity(outs[0][0], l_ref, what=cfg + " host losses")
        assert_parity(outs[0][1], g_ref, what=cfg + " host grads")


def test_fuzz_every_kernel_vs_oracle(e2e):
    """Random shapes, dtypes, strides, blanks, emission scales, tight alignments and infeasible rows, each on a
    randomly forced lattice kernel (or the library's own dispatch), fused and split: losses, gradients and
    NaN / inf positions against the oracle.  V = 2 is in the list (round 1 left it out)."""
    import random
    rng = random.Random(20261018)
    for it in range(140):
        B = rng.choice([1, 2, 3, 5, 8, 17, 80])
        T_ = rng.choice([1, 2, 3, 7, 8, 9, 31, 32, 33, 64, 100, 129, 257])
        V = rng.choice([2, 3, 5, 29, 32, 33, 64, 96, 128, 200])
        Lmax = rng.choice([0, 1, 2, 5, 31, 32, 62, 63, 64, 65, 127, 128, 200, 255, 256, 300])
        dt = rng.choice([torch.float32, torch.float32, torch.bfloat16, torch.float16, torch.float64])
        from_logits = rng.random() < 0.5
        time_major = rng.random() < 0.3
        fused = rng.random() < 0.7
        kind = rng.choice([AUTO, AUTO, GENERAL, WAVE, SWEEP])
        if V == 2 and kind == SWEEP:
            kind = AUTO                      # never dispatched there (see test_single_label_symbol_tight_peaky)
        blank = rng.randrange(V)
        scale = rng.choice([1.0, 1.0, 4.0, 10.0])
        g = torch.Generator().manual_seed(rng.randrange(1 << 30))
        x = torch.randn(B, T_, V, generator=g) * scale
        if not from_logits:
            x = torch.log_softmax(x, 2)
        x = x.to(dt)
        tl = torch.randint(0, Lmax + 1, (B,), generator=g)
        tg = torch.randint(0, V, (B, max(Lmax, 1)), generator=g)[:, :Lmax]
        tg = torch.where(tg == blank, (tg + 1) % V, tg)
        ll = torch.randint(1, T_ + 1, (B,), generator=g)
        mode = rng.random()
        if mode < 0.4:
            ll = torch.maximum(ll, torch.minimum(tl * 2, torch.tensor(T_)))
        elif mode < 0.6 and Lmax > 0:        # tight alignments: T_i = L_i + repeats (+0..2)
            rep = torch.tensor([int((tg[i, 1:tl[i]] == tg[i, :max(int(tl[i]) - 1, 0)]).sum()) if tl[i] > 1 else 0 for i in range(B)])
            ll = torch.clamp(tl + rep + torch.randint(0, 3, (B,), generator=g), 1, T_)
        xc = x.cuda()
        if time_major:
            xc = xc.permute(1, 0, 2).contiguous().permute(1, 0, 2)
        args = (xc, *cuda(tg, ll, tl))

        def run():
            eng = e2e.CTCLossEngine(blank)
            if fused:
                return eng.compute(*args, from_logits=from_logits)
            losses, st = eng.forward(*args, from_logits=from_logits)
            return losses, eng.backward(st)
        l_gpu, g_gpu = _with_kernel(kind, run)
        xr = x.double() if dt == torch.float64 else x.float()
        l_ref, g_ref = oracle.engine(blank).compute(torch.log_softmax(xr, 2) if from_logits else xr, tg, ll, tl)
        if from_logits:
            g_ref = _logits_grad_ref(g_ref, l_ref, ll)
        tol = 1e-5 if dt in (torch.float32, torch.float64) else (2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -10)
        what = "fuzz case %d (kernel %d fused %s B=%d T=%d V=%d Lmax=%d %s from_logits=%s tm=%s blank=%d x%g)" % (
            it, kind, fused, B, T_, V, Lmax, dt, from_logits, time_major, blank, scale)
        cast = (lambda t: t) if dt == torch.float64 else (lambda t: t.to(dt).float())
        assert_parity(l_gpu.double(), cast(l_ref).double(), rtol=tol, atol=tol, what=what + " losses")
        assert_parity(g_gpu.double(), cast(g_ref).double(), rtol=tol, atol=max(tol, 1e-5), what=what + " grads")
