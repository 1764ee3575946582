"""Differential fuzz: wave kernel (forced) vs the sweep / lattice kernels on random shapes."""
import os, sys, random
import faulthandler; faulthandler.dump_traceback_later(500, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from end2end_b200 import CTCLossEngine
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 150
bad = 0
for it in range(N):
    B = rng.choice([1, 2, 3, 5, 8, 17, 40])
    T = rng.choice([1, 2, 3, 7, 8, 9, 31, 32, 33, 64, 100, 129, 257, 400])
    V = rng.choice([2, 3, 5, 29, 32, 33, 64, 96, 128])
    Lmax = rng.choice([0, 1, 2, 5, 31, 32, 63, 64, 65, 127, 128, 200, 255])
    dt = rng.choice([torch.float32, torch.float32, torch.bfloat16, torch.float16])
    fl = rng.random() < 0.5
    tm = rng.random() < 0.3
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
        if V == 2: pass
    ll = torch.randint(1, T + 1, (B,), generator=g)
    if rng.random() < 0.5:
        ll = torch.maximum(ll, torch.minimum(tl * 2, torch.tensor(T)))
    xc = x.cuda()
    if tm:
        xc = xc.permute(1, 0, 2).contiguous().permute(1, 0, 2)
    args = (xc, tg.cuda(), ll.cuda(), tl.cuda())
    res = {}
    for wv in ("1", "0"):
        os.environ["E2E_CTC_WAVE"] = wv
        l, gr = CTCLossEngine(blank).compute(*args, from_logits=fl)
        res[wv] = (l.float().cpu(), gr.float().cpu())
    tol = 2e-5 if dt == torch.float32 else (2.0 ** -7 if dt == torch.bfloat16 else 2.0 ** -10)
    (l1, g1), (l0, g0) = res["1"], res["0"]
    ok = torch.equal(torch.isnan(l1), torch.isnan(l0)) and torch.equal(torch.isinf(l1), torch.isinf(l0)) and torch.equal(torch.isnan(g1), torch.isnan(g0))
    fin = torch.isfinite(l0)
    ok = ok and bool(((l1[fin] - l0[fin]).abs() <= tol + tol * l0[fin].abs()).all())
    gf = torch.isfinite(g0)
    ok = ok and bool(((g1[gf] - g0[gf]).abs() <= tol + tol * g0[gf].abs()).all())
    if not ok:
        import oracle
        lp = torch.log_softmax(x.float(), 2) if fl else x.float()
        lr, grr = oracle.engine(blank).compute(lp, tg, ll, tl)
        if fl:
            for r_, n_ in enumerate(ll.tolist()):
                grr[r_, n_:] = 0
                if not torch.isfinite(lr[r_]): grr[r_] = float("nan")
        gfin = torch.isfinite(grr)
        e1 = float((g1[gfin] - grr[gfin]).abs().max()); e0 = float((g0[gfin] - grr[gfin]).abs().max())
        d = (g1 - g0).abs(); d[~torch.isfinite(d)] = 0
        idx = (d > 0.5).nonzero()[:6].tolist()
        print("   vs oracle: wave err %.3e, sweep err %.3e; first big diffs (b,t,v): %s; ll %s tl %s" % (e1, e0, idx, [ll[i[0]].item() for i in idx[:3]], [tl[i[0]].item() for i in idx[:3]]), flush=True)
        bad += 1
        print("MISMATCH it %d B %d T %d V %d Lmax %d dt %s from_logits %s tm %s blank %d scale %g | loss err %.3e grad err %.3e nan_eq %s" % (
            it, B, T, V, Lmax, dt, fl, tm, blank, scale, float((l1[fin] - l0[fin]).abs().max()) if fin.any() else 0,
            float((g1[gf] - g0[gf]).abs().max()) if gf.any() else 0, torch.equal(torch.isnan(g1), torch.isnan(g0))), flush=True)
print("fuzz done: %d cases, %d mismatches" % (N, bad))
