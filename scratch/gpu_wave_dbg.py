import sys, os, time
import faulthandler; faulthandler.dump_traceback_later(250, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from end2end_b200 import CTCLossEngine
g = np.load('tests/golden/c2_b4_peaky.npz')
x = torch.from_numpy(g['x']); lp = torch.log_softmax(x, 2)
tg, ll, tl = [torch.from_numpy(g[k]) for k in ('targets', 'logits_lengths', 'targets_lengths')]
print("ll", ll.tolist(), "tl", tl.tolist())
eng = CTCLossEngine(0)
for rep in range(2):
    l, gr = eng.compute(lp.cuda(), tg.cuda(), ll.cuda(), tl.cuda())
    l = l.cpu(); gr = gr.cpu()
    ref = torch.from_numpy(g['engine_grads'])
    print("loss", l.tolist(), g['engine_losses'].tolist())
    err = (gr - ref).abs()
    bad = (err > 1e-5 + 1e-5 * ref.abs()).nonzero()
    print("nbad", len(bad))
    for b_, t_, v_ in bad[:60].tolist():
        print("  b %d t %d v %d ours % .6f ref % .6f" % (b_, t_, v_, gr[b_, t_, v_], ref[b_, t_, v_]))
# ---- timing: per-call event times ----
import oracle
for B in (64, 8, 1):
    xs, tgs, lls, tls = oracle.make_inputs(B, 400, 29, 100, 200, 1)
    xs = xs.cuda(); tgs = tgs.cuda(); lls = lls.cuda(); tls = tls.cuda()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(31)]
    for _ in range(3): eng.step(xs, tgs, lls, tls, from_logits=True, grad_scale=1.0 / B, reduce_scale=1.0 / B)
    torch.cuda.synchronize()
    evs[0].record()
    for i in range(30):
        eng.step(xs, tgs, lls, tls, from_logits=True, grad_scale=1.0 / B, reduce_scale=1.0 / B)
        evs[i + 1].record()
    torch.cuda.synchronize()
    print("B", B, "per-call us:", [int(evs[i].elapsed_time(evs[i + 1]) * 1000) for i in range(30)])
