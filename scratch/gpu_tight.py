import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle
from end2end_b200 import CTCLossEngine
g = torch.Generator().manual_seed(5)
B, V = 6, 5
tl = torch.tensor([55, 73, 66, 40, 120, 9])
T_ = 140
tg = torch.randint(1, V, (B, 120), generator=g)
ll = torch.zeros(B, dtype=torch.int64)
for b in range(B):
    L = int(tl[b]); rep = int((tg[b, 1:L] == tg[b, :L - 1]).sum()); ll[b] = L + rep + (b % 3)
x = torch.randn(B, T_, V, generator=g) * float(sys.argv[1] if len(sys.argv) > 1 else 10.0)
lp = torch.log_softmax(x, 2)
l_ref, g_ref = oracle.engine(0).compute(lp, tg, ll, tl)
print("ll", ll.tolist(), "tl", tl.tolist())
print("ref ", [round(v, 3) for v in l_ref.tolist()])
for wv in ("1", "0"):
    os.environ["E2E_CTC_WAVE"] = wv
    l, gr = CTCLossEngine(0).compute(lp.cuda(), tg.cuda(), ll.cuda(), tl.cuda())
    gerr = [(gr[b].cpu() - g_ref[b]).abs().max().item() for b in range(B)]
    print("wave" if wv == "1" else "sweep", [round(v, 3) for v in l.cpu().tolist()], "grad err per utt", ["%.1e" % e for e in gerr])
