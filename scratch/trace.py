"""Per-chunk pipeline timeline of the lattice kernel (E2E_CTC_TRACE=1)."""
import os, sys, ctypes
os.environ["E2E_CTC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from end2end_b200 import CTCLossEngine, _lib
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
B, T, V, Lmin, Lmax, seed, dt, full, desc = bench.WORKLOADS[wl]
x, tg, ll, tl = bench.make_inputs(B, T, V, Lmin, Lmax, seed, torch.float32, full)
eng = CTCLossEngine(0)
args = (x.cuda(), tg.cuda(), ll.cuda(), tl.cuda())
for _ in range(3):
    eng.step(*args, from_logits=True)
torch.cuda.synchronize()
L = _lib.load()
NCH, NEV = 128, 4
n = 2 * 2 * 3 * NCH * NEV
buf = (ctypes.c_longlong * n)()
L.e2e_ctc_debug_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
got = L.e2e_ctc_debug_trace_read(buf, n)
import numpy as np
a = np.array(buf[:], dtype=np.int64).reshape(2, 2, 3, NCH, NEV)
print("T0 =", int(ll[0]), "L0 =", int(tl[0]))
for d in (0, 1):
    tr = a[0, d]
    t0 = tr[tr > 0].min()
    print("sweep", "bwd" if d else "fwd", " (cycles relative to first stamp)")
    print(" chunk | prod: cpwait emptyok convdone | latt: waitstart fullok | comb: waitstart latdoneok done")
    for c in range(NCH):
        if tr[0, c, 0] == 0 and tr[1, c, 0] == 0:
            break
        r = lambda role, ev: int(tr[role, c, ev] - t0) if tr[role, c, ev] else -1
        print(" %4d | %8d %8d %8d | %8d %8d | %8d %8d %8d" % (c, r(0, 0), r(0, 1), r(0, 2), r(1, 0), r(1, 1), r(2, 0), r(2, 1), r(2, 2)))
