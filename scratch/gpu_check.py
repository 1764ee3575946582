import sys, time, os
import faulthandler; faulthandler.dump_traceback_later(400, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle
from end2end_b200 import CTCLoss, CTCDecoder, CTCLossEngine

def cmp(name, a, b, rtol=1e-5, atol=1e-5):
    a = a.detach().cpu().double(); b = b.detach().cpu().double()
    nan_eq = torch.equal(torch.isnan(a), torch.isnan(b)); inf_eq = torch.equal(torch.isinf(a), torch.isinf(b))
    fin = torch.isfinite(a) & torch.isfinite(b)
    d = (a[fin] - b[fin]).abs()
    viol = (d > atol + rtol * b[fin].abs()).sum().item()
    print("%-28s maxabs %.3e viol %d/%d nan_eq %s inf_eq %s" % (name, d.max().item() if d.numel() else 0, viol, d.numel(), nan_eq, inf_eq), flush=True)
    return viol == 0 and nan_eq and inf_eq

ok = True
for cfg in ["c1", "c2", "c4", "c3"]:
    B,T,V,Lmin,Lmax,seed,dtype,full = oracle.CONFIGS[cfg]
    if cfg == "c3": B = 128
    x, tg, ll, tl = oracle.make_inputs(B,T,V,Lmin,Lmax,seed,dtype=torch.float32, full_length=full)
    for after in (False, True):
        xin = torch.log_softmax(x, 2) if after else x
        eng = oracle.engine(0)
        xr = xin.clone().requires_grad_()
        t0=time.time(); lref = oracle.ctc_loss_module(eng, xr, tg, ll, tl, reduce=True, size_average=True, after_logsoftmax=after); lref.backward(); tr=time.time()-t0
        xg = xin.cuda().requires_grad_()
        crit = CTCLoss(reduce=True, size_average=True, after_logsoftmax=after)
        l = crit(xg, tg.cuda(), ll.cuda(), tl.cuda()); l.backward(); torch.cuda.synchronize()
        print(cfg, "after" if after else "logits", "loss", l.item(), lref.item(), "ref time %.3f" % tr)
        ok &= cmp(cfg+" loss", l, lref)
        ok &= cmp(cfg+" grad", xg.grad, xr.grad, atol=1e-5/B*1 if False else 1e-5)
    # per-utterance + engine.compute contract
    lp = torch.log_softmax(x, 2)
    l1, g1 = oracle.engine(0).compute(lp, tg, ll, tl)
    l2, g2 = CTCLossEngine(0).compute(lp.cuda(), tg.cuda(), ll.cuda(), tl.cuda())
    ok &= cmp(cfg+" engine loss", l2, l1); ok &= cmp(cfg+" engine grads", g2, g1)
    # host path
    l3, g3 = CTCLossEngine(0).compute(lp, tg, ll, tl)
    ok &= cmp(cfg+" host loss", l3, l1); ok &= cmp(cfg+" host grads", g3, g1)
    # greedy
    r = CTCDecoder(beam_width=1).decode(x.cuda(), ll.cuda())
    o = oracle.greedy_decode(x, ll)
    eq = torch.equal(r.decoded_targets, o[0]) and torch.equal(r.decoded_targets_lengths, o[1])
    print(cfg, "greedy equal", eq); ok &= eq
    # timing
    crit = CTCLoss(reduce=True, size_average=True)
    xg = x.cuda().requires_grad_(); tgc, llc, tlc = tg.cuda(), ll.cuda(), tl.cuda()
    for _ in range(3):
        xg.grad=None; crit(xg, tgc, llc, tlc).backward()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        xg.grad=None; crit(xg, tgc, llc, tlc).backward()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/20
    print(cfg, "fwd+bwd %.3f ms -> %.0f utt/s" % (ms, B/ms*1e3), flush=True)
# edge cases
eng_o = oracle.engine(0); eng = CTCLossEngine(0)
lp = torch.log_softmax(torch.randn(6, 7, 5, generator=torch.Generator().manual_seed(5)), 2)
tg = torch.tensor([[1,1,2],[1,2,3],[2,2,2],[4,0,0],[1,2,1],[3,3,1]]); tl = torch.tensor([3,3,3,1,0,2]); ll = torch.tensor([7,3,4,1,5,2])
l1, g1 = eng_o.compute(lp, tg, ll, tl); l2, g2 = eng.compute(lp.cuda(), tg.cuda(), ll.cuda(), tl.cuda())
print(l1, l2.cpu())
ok &= cmp("edge loss", l2, l1); ok &= cmp("edge grads", g2, g1)
print("ALL OK" if ok else "FAILURES")
