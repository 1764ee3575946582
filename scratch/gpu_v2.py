import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle
from end2end_b200 import CTCLossEngine
g = torch.Generator().manual_seed(3)
B, T_, V, Lmax, blank = 17, 257, 2, 255, 1
x = torch.log_softmax(torch.randn(B, T_, V, generator=g) * 10.0, 2)
tl = torch.randint(0, Lmax + 1, (B,), generator=g)
tl[:6] = torch.tensor([0, 1, 2, 60, 100, 129])
tg = torch.zeros(B, Lmax, dtype=torch.int64)
ll = torch.randint(1, T_ + 1, (B,), generator=g)
ll[:6] = torch.tensor([5, 9, 20, 200, 257, 257])
l_ref, g_ref = oracle.engine(blank).compute(x, tg, ll, tl)
print("ll", ll.tolist()); print("tl", tl.tolist())
print("ref  ", ["%.2f" % v for v in l_ref.tolist()])
for wv in ("1", "0"):
    os.environ["E2E_CTC_WAVE"] = wv
    l, gr = CTCLossEngine(blank).compute(x.cuda(), tg.cuda(), ll.cuda(), tl.cuda())
    gr = gr.cpu(); l = l.cpu()
    errs = []
    for b in range(B):
        fin = torch.isfinite(g_ref[b])
        errs.append(float((gr[b][fin] - g_ref[b][fin]).abs().max()) if fin.any() else -1)
    print("wave " if wv == "1" else "sweep", ["%.2f" % v for v in l.tolist()])
    print("   grad err/utt", ["%.0e" % e for e in errs], "nan_eq", torch.equal(torch.isnan(gr), torch.isnan(g_ref)))
