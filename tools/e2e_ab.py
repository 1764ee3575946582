"""A/B of the host-buffer call's result allocation on one box: python tools/e2e_ab.py [cfg]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle
from end2end_b200 import CTCLossEngine
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
B, T, V, Lmin, Lmax, seed, dtype, full = oracle.CONFIGS[cfg]
x, tg, ll, tl = oracle.make_inputs(B, T, V, Lmin, Lmax, seed, dtype=dtype, full_length=full)
eng = CTCLossEngine(0); hx, htg, hll, htl = (t.pin_memory() for t in (x, tg, ll, tl))
for rnd in range(2):
    for pool in (True, False):
        if pool: os.environ.pop("E2E_CTC_NO_PINNED_POOL", None)
        else: os.environ["E2E_CTC_NO_PINNED_POOL"] = "1"
        for _ in range(5): eng.compute(hx, htg, hll, htl, from_logits=True)
        for hold in (False, True):
            t0 = time.perf_counter()
            for _ in range(100):
                if hold:
                    hl, hg = eng.compute(hx, htg, hll, htl, from_logits=True); float(hl[0])
                else:
                    float(eng.compute(hx, htg, hll, htl, from_logits=True)[0][0])
            dt = (time.perf_counter() - t0) / 100
            print("pool" if pool else "torch", "hold" if hold else "drop", "%.1f us/call -> %.0f utt/s" % (dt * 1e6, B / dt))
            hl = hg = None
