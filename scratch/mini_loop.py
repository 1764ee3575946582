import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from end2end_b200 import CTCLoss
import bench
B, T, V, Lmin, Lmax, seed, dt, full, desc = bench.WORKLOADS["c2"]
nb = 89
bs = []
for i in range(nb):
    x, tg, ll, tl = bench.make_inputs(B, T, V, Lmin, Lmax, seed + i, torch.float32, full)
    bs.append((x.cuda().requires_grad_(), tg.cuda(), ll.cuda(), tl.cuda()))
crit = CTCLoss(reduce=True, size_average=True)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 120):
    x, tg, ll, tl = bs[i % nb]
    x.grad = None
    crit(x, tg, ll, tl).backward()
torch.cuda.synchronize()
