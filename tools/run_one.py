"""Run one workload's training step a few times (for ncu / sanitizer captures): python tools/run_one.py c2 [steps] [batch]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import oracle  # noqa: E402
from end2end_b200 import CTCDecoder, CTCLossEngine  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
B, T, V, Lmin, Lmax, seed, dtype, full = oracle.CONFIGS[cfg]
if len(sys.argv) > 3:
    B = int(sys.argv[3])
x, tg, ll, tl = oracle.make_inputs(B, T, V, Lmin, Lmax, seed, dtype=dtype, full_length=full)
eng = CTCLossEngine(0)
xg, tgc, llc, tlc = x.cuda(), tg.cuda(), ll.cuda(), tl.cuda()
for _ in range(steps):
    losses, grads, red, _ = eng.step(xg, tgc, llc, tlc, True, 1.0 / B, 1.0 / B)
if "--greedy" in sys.argv:
    CTCDecoder(beam_width=1).decode(xg, llc)
torch.cuda.synchronize()
print(cfg, "loss", float(red))
