"""Where does the host time of one CTCLoss fwd+bwd step go? (run on the GPU box)"""
import cProfile, pstats, sys, os, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from end2end_b200 import CTCLoss
import bench
B, T, V, Lmin, Lmax, seed, dt, full, desc = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c1"]
x, tg, ll, tl = bench.make_inputs(B, T, V, Lmin, Lmax, seed, torch.float32, full)
x = x.cuda().requires_grad_(); tg = tg.cuda(); ll = ll.cuda(); tl = tl.cuda()
crit = CTCLoss(reduce=True, size_average=True)
def step():
    x.grad = None
    loss = crit(x, tg, ll, tl)
    loss.backward()
for _ in range(20): step()
torch.cuda.synchronize()
N = 500
t0 = time.perf_counter()
for _ in range(N): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host issue time per step %.1f us; with final sync %.1f us" % ((t1 - t0) / N * 1e6, (t2 - t0) / N * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(N): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
