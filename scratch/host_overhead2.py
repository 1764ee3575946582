"""Host time of the module step with rotating batches (bench.py's timed loop)."""
import cProfile, pstats, sys, os, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from end2end_b200 import CTCLoss
import bench
B, T, V, Lmin, Lmax, seed, dt, full, desc = bench.WORKLOADS["c2"]
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 89
bs = []
for i in range(nb):
    x, tg, ll, tl = bench.make_inputs(B, T, V, Lmin, Lmax, seed + i, torch.float32, full)
    bs.append((x.cuda().requires_grad_(), tg.cuda(), ll.cuda(), tl.cuda()))
crit = CTCLoss(reduce=True, size_average=True)
def step(i):
    x, tg, ll, tl = bs[i % nb]
    x.grad = None
    loss = crit(x, tg, ll, tl)
    loss.backward()
for i in range(20): step(i)
torch.cuda.synchronize()
N = 300
t0 = time.perf_counter()
for i in range(N): step(i)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("nb %d: host issue time per step %.1f us; with final sync %.1f us" % (nb, (t1 - t0) / N * 1e6, (t2 - t0) / N * 1e6))
print(torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_stats()["num_device_alloc"], torch.cuda.memory_stats()["num_device_free"])
pr = cProfile.Profile(); pr.enable()
for i in range(N): step(i)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(12); print(s.getvalue()[:3000])
print(torch.cuda.memory_stats()["num_device_alloc"], torch.cuda.memory_stats()["num_device_free"])
