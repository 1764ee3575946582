import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle
from end2end_b200 import CTCLossEngine, CTCLoss
eng = CTCLossEngine(0)
B = 64
def mk(seed):
    xs, tgs, lls, tls = oracle.make_inputs(B, 400, 29, 100, 200, seed)
    return xs.cuda(), tgs.cuda(), lls.cuda(), tls.cuda()
def timeit(fn, n=60):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(n): fn(i)
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000, (t1 - t0) / n * 1e6
for nb in (1, 4, 30, 90):
    bs = [mk(100 + i) for i in range(nb)]
    f = lambda i: eng.step(*bs[i % nb], from_logits=True, grad_scale=1.0 / B, reduce_scale=1.0 / B)
    print("eng.step  nbatches %3d: gpu %.1f us/step, host issue %.1f us/step" % ((nb,) + timeit(f)), flush=True)
crit = CTCLoss(reduce=True, size_average=True)
for nb in (1, 90):
    bs = [mk(100 + i) for i in range(nb)]
    xs = [b[0].requires_grad_() for b in bs]
    def g(i):
        x = xs[i % nb]; x.grad = None
        crit(x, *bs[i % nb][1:]).backward()
    print("module    nbatches %3d: gpu %.1f us/step, host issue %.1f us/step" % ((nb,) + timeit(g)), flush=True)
