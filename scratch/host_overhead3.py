"""Host-tensor engine call: where does the time go?"""
import cProfile, pstats, sys, os, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from end2end_b200 import CTCLossEngine
import bench
B, T, V, Lmin, Lmax, seed, dt, full, desc = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
hx, htg, hll, htl = bench.make_inputs(B, T, V, Lmin, Lmax, seed, torch.float32, full)
hx = hx.pin_memory()
eng = CTCLossEngine(0)
for _ in range(5): eng.compute(hx, htg, hll, htl, from_logits=True)
N = 100
t0 = time.perf_counter()
for _ in range(N): eng.compute(hx, htg, hll, htl, from_logits=True)
t1 = time.perf_counter()
print("per call %.1f us" % ((t1 - t0) / N * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(N): eng.compute(hx, htg, hll, htl, from_logits=True)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(10); print(s.getvalue()[:2500])
