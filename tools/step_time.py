"""Step time of one workload through engine.step (kernel-level experiments): python tools/step_time.py cfg [batch] [reps] [Lmin Lmax]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle
from end2end_b200 import CTCLossEngine, _lib
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
B, T, V, Lmin, Lmax, seed, dtype, full = oracle.CONFIGS[cfg]
if len(sys.argv) > 2 and int(sys.argv[2]) > 0: B = int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
if len(sys.argv) > 5:       # override the target-length range (dispatch experiments)
    Lmin, Lmax = int(sys.argv[4]), int(sys.argv[5])
x, tg, ll, tl = oracle.make_inputs(B, T, V, Lmin, Lmax, seed, dtype=dtype, full_length=full)
if os.environ.get("FORCE"): _lib.force_kernel(int(os.environ["FORCE"]))
eng = CTCLossEngine(0)
xs = [x.cuda(), torch.randn(B, T, V, device="cuda").to(dtype)]
tgc, llc, tlc = tg.cuda(), ll.cuda(), tl.cuda()
for i in range(3): eng.step(xs[i % 2], tgc, llc, tlc, True, 1.0 / B, 1.0 / B)
torch.cuda.synchronize()
_lib.profile_enable(True); _lib.profile_read()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps): out = eng.step(xs[i % 2], tgc, llc, tlc, True, 1.0 / B, 1.0 / B)
e1.record(); torch.cuda.synchronize()
prof = _lib.profile_read(); _lib.profile_enable(False)
print(cfg, "B", B, "step %.1f us" % (e0.elapsed_time(e1) / reps * 1e3), {k: round(v[0] / v[1] * 1e3, 1) for k, v in prof.items() if v[1]}, "loss", float(out[2]), os.environ.get("TAG", ""))
