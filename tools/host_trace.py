"""Device/host timeline of the chunked host-buffer call (E2E_CTC_HOST_TRACE=1): python tools/host_trace.py [cfg] [chunks]"""
import os, sys
os.environ["E2E_CTC_HOST_TRACE"] = "1"
if len(sys.argv) > 2: os.environ["E2E_CTC_HOST_CHUNKS"] = sys.argv[2]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, torch, oracle
from end2end_b200 import CTCLossEngine
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
B, T, V, Lmin, Lmax, seed, dtype, full = oracle.CONFIGS[cfg]
x, tg, ll, tl = oracle.make_inputs(B, T, V, Lmin, Lmax, seed, dtype=dtype, full_length=full)
eng = CTCLossEngine(0)
xp = x.pin_memory()
for i in range(8):
    t0 = time.perf_counter()
    l, g = eng.compute(xp, tg, ll, tl, from_logits=True)
    sys.stderr.write("call %d: %.0f us in Python\n" % (i, (time.perf_counter() - t0) * 1e6))
