"""Developer harness (GPU box): the CUDA path vs the CPU oracle on the BASELINE configs, an edge-case batch and a
seeded random-shape fuzz, plus per-kernel device times.  `python tools/gpu_check.py [configs|fuzz|time|all] [seed] [n]`."""
import os
import random
import sys
import time
import faulthandler

faulthandler.dump_traceback_later(900, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import oracle  # noqa: E402
from end2end_b200 import CTCDecoder, CTCLoss, CTCLossEngine, _lib  # noqa: E402


def cmp(name, a, b, rtol=1e-5, atol=1e-5, quiet=False):
    a = a.detach().cpu().double(); b = b.detach().cpu().double()
    nan_eq = torch.equal(torch.isnan(a), torch.isnan(b)); inf_eq = torch.equal(torch.isinf(a), torch.isinf(b))
    fin = torch.isfinite(a) & torch.isfinite(b)
    d = (a[fin] - b[fin]).abs()
    viol = int((d > atol + rtol * b[fin].abs()).sum())
    ok = viol == 0 and nan_eq and inf_eq
    if not quiet or not ok:
        print("%-30s maxabs %.3e viol %d/%d nan_eq %s inf_eq %s" % (name, float(d.max()) if d.numel() else 0, viol, d.numel(), nan_eq, inf_eq), flush=True)
    return ok


def oracle_ref(x, tg, ll, tl, blank, from_logits):
    """(losses, grads) the reference would return, with the fused-logits conventions applied."""
    lp = torch.log_softmax(x.float(), 2) if from_logits else x.float()
    lr, gr = oracle.engine(blank).compute(lp, tg, ll, tl)
    if from_logits:
        for r_, n_ in enumerate(ll.tolist()):
            gr[r_, n_:] = 0
            if not torch.isfinite(lr[r_]):
                gr[r_] = float("nan")
    return lr, gr


def run_configs():
    ok = True
    for cfg in ["c1", "c2", "c4", "c3", "c5"]:
        B, T, V, Lmin, Lmax, seed, dtype, full = oracle.CONFIGS[cfg]
        if cfg == "c3": B = 128
        if cfg == "c5": B = 6
        x, tg, ll, tl = oracle.make_inputs(B, T, V, Lmin, Lmax, seed, dtype=torch.float32, full_length=full)
        for after in (False, True):
            xin = torch.log_softmax(x, 2) if after else x
            eng = oracle.engine(0)
            xr = xin.clone().requires_grad_()
            t0 = time.time()
            lref = oracle.ctc_loss_module(eng, xr, tg, ll, tl, reduce=True, size_average=True, after_logsoftmax=after); lref.backward()
            tr = time.time() - t0
            xg = xin.cuda().requires_grad_()
            crit = CTCLoss(reduce=True, size_average=True, after_logsoftmax=after)
            l = crit(xg, tg.cuda(), ll.cuda(), tl.cuda()); l.backward(); torch.cuda.synchronize()
            print(cfg, "after" if after else "logits", "loss", l.item(), lref.item(), "ref time %.3f" % tr, flush=True)
            ok &= cmp(cfg + " loss", l, lref)
            ok &= cmp(cfg + " grad", xg.grad, xr.grad)
        lp = torch.log_softmax(x, 2)
        l1, g1 = oracle.engine(0).compute(lp, tg, ll, tl)
        l2, g2 = CTCLossEngine(0).compute(lp.cuda(), tg.cuda(), ll.cuda(), tl.cuda())
        ok &= cmp(cfg + " engine loss", l2, l1); ok &= cmp(cfg + " engine grads", g2, g1)
        # split forward / backward (gather-mode lattice + K1 + K3 for every alphabet)
        e = CTCLossEngine(0)
        lf, st = e.forward(lp.cuda(), tg.cuda(), ll.cuda(), tl.cuda())
        gb = e.backward(st)
        ok &= cmp(cfg + " split loss", lf, l1); ok &= cmp(cfg + " split grads", gb, g1)
        l3, g3 = CTCLossEngine(0).compute(lp, tg, ll, tl)
        ok &= cmp(cfg + " host loss", l3, l1); ok &= cmp(cfg + " host grads", g3, g1)
        r = CTCDecoder(beam_width=1).decode(x.cuda(), ll.cuda())
        o = oracle.greedy_decode(x, ll)
        eq = torch.equal(r.decoded_targets, o[0]) and torch.equal(r.decoded_targets_lengths, o[1])
        print(cfg, "greedy equal", eq); ok &= eq
    eng_o = oracle.engine(0); eng = CTCLossEngine(0)
    lp = torch.log_softmax(torch.randn(6, 7, 5, generator=torch.Generator().manual_seed(5)), 2)
    tg = torch.tensor([[1, 1, 2], [1, 2, 3], [2, 2, 2], [4, 0, 0], [1, 2, 1], [3, 3, 1]]); tl = torch.tensor([3, 3, 3, 1, 0, 2]); ll = torch.tensor([7, 3, 4, 1, 5, 2])
    l1, g1 = eng_o.compute(lp, tg, ll, tl); l2, g2 = eng.compute(lp.cuda(), tg.cuda(), ll.cuda(), tl.cuda())
    print(l1, l2.cpu())
    ok &= cmp("edge loss", l2, l1); ok &= cmp("edge grads", g2, g1)
    # -inf log-probs / logits (masked symbols)
    x = torch.randn(3, 20, 6, generator=torch.Generator().manual_seed(9)); x[:, :, 4] = float("-inf")
    tg = torch.tensor([[1, 2, 3], [1, 1, 2], [4, 2, 0]]); tl = torch.tensor([3, 3, 2]); ll = torch.tensor([20, 15, 9])
    for fl in (True, False):
        xin = x if fl else torch.log_softmax(x, 2)
        lr, gr = oracle_ref(xin, tg, ll, tl, 0, fl)
        l2, g2 = eng.compute(xin.cuda(), tg.cuda(), ll.cuda(), tl.cuda(), from_logits=fl)
        ok &= cmp("-inf fl=%s loss" % fl, l2, lr); ok &= cmp("-inf fl=%s grads" % fl, g2, gr)
    print("CONFIGS OK" if ok else "CONFIGS FAILURES", flush=True)
    return ok


def run_fuzz(seed, N):
    rng = random.Random(seed)
    bad = 0
    for it in range(N):
        B = rng.choice([1, 2, 3, 5, 8, 17, 40, 90])
        T = rng.choice([1, 2, 3, 7, 8, 9, 31, 32, 33, 64, 100, 129, 257, 400])
        V = rng.choice([2, 3, 5, 29, 32, 33, 64, 96, 128, 200])
        Lmax = rng.choice([0, 1, 2, 5, 31, 32, 62, 63, 64, 65, 126, 127, 128, 200, 255, 256, 300])
        dt = rng.choice([torch.float32, torch.float32, torch.bfloat16, torch.float16, torch.float64])
        fl = rng.random() < 0.5
        tm = rng.random() < 0.3
        fused = rng.random() < 0.7
        kern = rng.choice([-1, -1, 0, 1, 2])
        if V == 2 and kern == 2:
            kern = 0      # the sweep kernel is never dispatched for a single label symbol (its open corner)
        blank = rng.randrange(V)
        scale = rng.choice([1.0, 1.0, 4.0, 10.0])
        g = torch.Generator().manual_seed(rng.randrange(1 << 30))
        x = (torch.randn(B, T, V, generator=g) * scale)
        if not fl:
            x = torch.log_softmax(x, 2)
        x = x.to(dt)
        tl = torch.randint(0, Lmax + 1, (B,), generator=g)
        tg = torch.randint(0, V, (B, max(Lmax, 1)), generator=g)[:, :Lmax] if Lmax > 0 else torch.zeros(B, 0, dtype=torch.int64)
        if Lmax > 0 and V > 1:
            tg = torch.where(tg == blank, (tg + 1) % V, tg)
        ll = torch.randint(1, T + 1, (B,), generator=g)
        mode = rng.random()
        if mode < 0.4:
            ll = torch.maximum(ll, torch.minimum(tl * 2, torch.tensor(T)))
        elif mode < 0.6 and Lmax > 0:     # tight alignments: T_i = L_i + repeats (+0..2)
            rep = torch.tensor([int((tg[i, 1:tl[i]] == tg[i, :max(int(tl[i]) - 1, 0)]).sum()) if tl[i] > 1 else 0 for i in range(B)])
            ll = torch.clamp(tl + rep + torch.randint(0, 3, (B,), generator=g), 1, T)
        xc = x.cuda()
        if tm:
            xc = xc.permute(1, 0, 2).contiguous().permute(1, 0, 2)
        args = (xc, tg.cuda(), ll.cuda(), tl.cuda())
        _lib.force_kernel(kern)
        eng = CTCLossEngine(blank)
        if fused:
            l, gr = eng.compute(*args, from_logits=fl)
        else:
            l, st = eng.forward(*args, from_logits=fl)
            gr = eng.backward(st)
        lr, grr = oracle_ref(x.double() if dt == torch.float64 else x, tg, ll, tl, blank, fl)
        tol = 1e-5 if dt in (torch.float32, torch.float64) else (2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -10)
        ok = cmp("loss", l.float(), lr.to(dt).float() if dt != torch.float64 else lr, tol, tol, quiet=True)
        ok &= cmp("grad", gr.float(), grr.to(dt).float() if dt != torch.float64 else grr, tol, max(tol, 1e-5), quiet=True)
        if not ok:
            bad += 1
            print("MISMATCH it %d kernel %d B %d T %d V %d Lmax %d dt %s from_logits %s tm %s fused %s blank %d scale %g ll %s tl %s" % (
                it, kern, B, T, V, Lmax, dt, fl, tm, fused, blank, scale, ll.tolist()[:6], tl.tolist()[:6]), flush=True)
    _lib.force_kernel(-1)
    print("fuzz seed %d: %d cases, %d mismatches" % (seed, N, bad), flush=True)
    return bad == 0


def run_time():
    for cfg in ["c1", "c2", "c3", "c4", "c5"]:
        B, T, V, Lmin, Lmax, seed, dtype, full = oracle.CONFIGS[cfg]
        if cfg == "c5": B = 512
        x, tg, ll, tl = oracle.make_inputs(B, T, V, Lmin, Lmax, seed, dtype=dtype, full_length=full)
        eng = CTCLossEngine(0)
        xg, tgc, llc, tlc = x.cuda(), tg.cuda(), ll.cuda(), tl.cuda()
        for _ in range(3):
            eng.step(xg, tgc, llc, tlc, True, 1.0 / B, 1.0 / B)
        torch.cuda.synchronize()
        _lib.profile_enable(True); _lib.profile_read()
        n = 10 if cfg != "c5" else 3
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            eng.step(xg, tgc, llc, tlc, True, 1.0 / B, 1.0 / B)
        e1.record(); torch.cuda.synchronize()
        prof = _lib.profile_read(); _lib.profile_enable(False)
        ms = e0.elapsed_time(e1) / n
        kern = {k: round(v[0] / v[1] * 1e3, 1) for k, v in prof.items() if v[1]}
        esz = 2 if dtype == torch.bfloat16 else 4
        print("%s B=%d step %.3f ms (with event hooks) -> %.0f utt/s; kernels (us): %s; step roofline frac %.4f" % (
            cfg, B, ms, B / ms * 1e3, kern, B * T * V * 2 * esz / (ms * 1e-3) / 6.5565e12), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 150
    ok = True
    if what in ("configs", "all"):
        ok &= run_configs()
    if what in ("fuzz", "all"):
        ok &= run_fuzz(seed, n)
    if what in ("time", "all"):
        run_time()
    print("ALL OK" if ok else "FAILURES")
