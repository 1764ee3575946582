"""Times the prefix-beam-search kernel on a BASELINE shape (device tensors, CUDA events) -- used for ncu captures.

    python tools/beam_probe.py [c2|c4|c1|c3] [beam_width] [reps]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (inputs only: the seeded draw of SURVEY 8(d))
from end2end_b200.engine import CTCBeamEngine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
beam = int(sys.argv[2]) if len(sys.argv) > 2 else 100
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
x, _, ll, _ = oracle.make_inputs(*oracle.CONFIGS[name][:7])
x, ll = x.cuda(), ll.cuda()
eng = CTCBeamEngine(0, beam)
for from_logits in (True,):
    eng.decode_device(x, ll, from_logits=from_logits)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dec, n, ties = eng.decode_device(x, ll, from_logits=from_logits)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("%s beam %d from_logits %d: %.3f ms (min of %d) -> %.0f utt/s; mean length %.1f; ties %d" % (
        name, beam, from_logits, min(ts), reps, x.size(0) / min(ts) * 1e3, float(n.float().mean()), int(ties.sum())))
