"""Cycle counters of the general lattice kernel (a -DFZ_DBG build): python tools/fz_dbg.py [cfg] [batch]
Run with E2E_CTC_LIB=end2end_b200/lib/libe2e_ctc_dbg.so; E2E_CTC_DBG=4 idles the helper warps' work."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle
from end2end_b200 import CTCLossEngine, _lib
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
B, T, V, Lmin, Lmax, seed, dtype, full = oracle.CONFIGS[cfg]
if len(sys.argv) > 2: B = int(sys.argv[2])
x, tg, ll, tl = oracle.make_inputs(B, T, V, Lmin, Lmax, seed, dtype=dtype, full_length=full)
_lib.force_kernel(int(os.environ.get("FORCE", "0")))
eng = CTCLossEngine(0)
xc, tgc, llc, tlc = x.cuda(), tg.cuda(), ll.cuda(), tl.cuda()
for _ in range(3):
    eng.step(xc, tgc, llc, tlc, True, 1.0 / B, 1.0 / B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): eng.step(xc, tgc, llc, tlc, True, 1.0 / B, 1.0 / B)
e1.record(); torch.cuda.synchronize()
print(cfg, "B", B, "step %.1f us" % (e0.elapsed_time(e1) / 5 * 1e3), "ll[0:2]", ll[:2].tolist(), "tl[0:2]", tl[:2].tolist())
