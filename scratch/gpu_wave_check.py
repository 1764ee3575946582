import sys, os, time
import faulthandler; faulthandler.dump_traceback_later(150, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle
from end2end_b200 import CTCLossEngine

def cmp(name, a, b, rtol=1e-5, atol=1e-5):
    a = a.detach().cpu().double(); b = b.detach().cpu().double()
    nan_eq = torch.equal(torch.isnan(a), torch.isnan(b)); inf_eq = torch.equal(torch.isinf(a), torch.isinf(b))
    fin = torch.isfinite(a) & torch.isfinite(b)
    d = (a[fin] - b[fin]).abs()
    viol = (d > atol + rtol * b[fin].abs()).sum().item()
    print("%-34s maxabs %.3e viol %d/%d nan_eq %s inf_eq %s" % (name, d.max().item() if d.numel() else 0, viol, d.numel(), nan_eq, inf_eq), flush=True)
    return viol == 0 and nan_eq and inf_eq

ok = True
cases = [("tiny", 2, 6, 5, 1, 2, 3), ("c1", 4, 50, 28, 10, 29, 0), ("nw2", 3, 150, 12, 70, 120, 5), ("c2b8", 8, 400, 29, 100, 200, 1),
         ("v96", 16, 128, 96, 20, 40, 2), ("nw8", 2, 700, 9, 300, 500, 6)]
only = sys.argv[1:] 
for name, B, T, V, Lmin, Lmax, seed in cases:
    if only and name not in only: continue
    x, tg, ll, tl = oracle.make_inputs(B, T, V, Lmin, Lmax, seed)
    lp = torch.log_softmax(x, 2)
    l1, g1 = oracle.engine(0).compute(lp, tg, ll, tl)
    for fl, inp in ((False, lp), (True, x)):
        t0 = time.time()
        l2, g2 = CTCLossEngine(0).compute(inp.cuda(), tg.cuda(), ll.cuda(), tl.cuda(), from_logits=fl)
        torch.cuda.synchronize()
        print(name, "from_logits", fl, "%.3fs" % (time.time() - t0), flush=True)
        ok &= cmp(name + " loss", l2, l1)
        gexp = g1.clone()
        if fl:
            for r, n in enumerate(ll.tolist()):
                gexp[r, n:] = 0
                if not torch.isfinite(l1[r]): gexp[r] = float("nan")
        ok &= cmp(name + " grads", g2, gexp)
print("ALL OK" if ok else "FAILED")
