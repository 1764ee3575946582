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
if len(sys.argv) > 3 and sys.argv[3].isdigit() and int(sys.argv[3]) > 0:
    B = int(sys.argv[3])
x, tg, ll, tl = oracle.make_inputs(B, T, V, Lmin, Lmax, seed, dtype=dtype, full_length=full)
eng = CTCLossEngine(0)
xg, tgc, llc, tlc = x.cuda(), tg.cuda(), ll.cuda(), tl.cuda()
for _ in range(steps):
    losses, grads, red, _ = eng.step(xg, tgc, llc, tlc, True, 1.0 / B, 1.0 / B)
if "--greedy" in sys.argv:
    for _ in range(2):
        CTCDecoder(beam_width=1).decode(xg, llc)
if "--scale" in sys.argv:        # an upstream gradient that is not 1: the in-place row scaling kernel does real work
    for _ in range(2):
        eng.scale_rows_(grads, torch.full((B,), 0.5, device="cuda", dtype=grads.dtype))
if "--noblank" in sys.argv:
    from end2end_b200.functions.ctc_without_blank import ctc_without_blank_3d_loss
    for _ in range(2):
        ctc_without_blank_3d_loss(torch.log_softmax(xg.float(), 2), tgc, llc, tlc)
if "--align" in sys.argv:
    from end2end_b200.utils.alignment import get_alignment_3d_device
    lp = torch.log_softmax(xg.float(), 2)
    for _ in range(2):
        get_alignment_3d_device(lp, tgc, llc, tlc)
torch.cuda.synchronize()
print(cfg, "loss", float(red))
