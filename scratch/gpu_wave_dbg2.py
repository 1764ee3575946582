import sys, os
import faulthandler; faulthandler.dump_traceback_later(250, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle
from end2end_b200 import CTCLossEngine
eng = CTCLossEngine(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
xs, tgs, lls, tls = oracle.make_inputs(B, 400, 29, 100, 200, 1, full_length=True)
xs = xs.cuda(); tgs = tgs.cuda(); lls = lls.cuda(); tls = tls.cuda()
for _ in range(2):
    eng.step(xs, tgs, lls, tls, from_logits=True, grad_scale=1.0 / B, reduce_scale=1.0 / B)
    torch.cuda.synchronize()
    sys.stderr.write("----\n")
