import sys, os, faulthandler, time
faulthandler.dump_traceback_later(90, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("OMP_NUM_THREADS", "8")
import torch, ctypes
print("torch ok", flush=True)
import oracle
from end2end_b200 import _lib
from end2end_b200.engine import CTCLossEngine, _Problem, _ptr, _stream
L = _lib.load()
x, tg, ll, tl = oracle.make_inputs(4, 50, 28, 10, 29, 0, full_length=True)
xg = x.cuda(); tgc, llc, tlc = tg.cuda(), ll.cuda(), tl.cuda()
torch.cuda.synchronize(); print("inputs on gpu", flush=True)
eng = CTCLossEngine(0)
pb = _Problem(0, xg, tgc, llc, tlc, True, xg.device)
ws = eng._workspace(pb); print("ws bytes", ws.numel(), flush=True)
losses = torch.zeros(4, device="cuda")
rc = L.e2e_ctc_loss_forward_device(ctypes.byref(pb.desc), _ptr(pb.logits), _ptr(pb.targets), _ptr(pb.logits_lengths), _ptr(pb.targets_lengths), _ptr(losses), _ptr(ws), ws.numel(), _stream(xg.device))
print("forward rc", rc, L.e2e_last_error_string(), flush=True)
torch.cuda.synchronize(); print("forward done", losses.cpu(), flush=True)
lref, gref = oracle.engine(0).compute(torch.log_softmax(x, 2), tg, ll, tl)
print("ref", lref, flush=True)
grads = torch.empty_like(xg)
rc = L.e2e_ctc_loss_backward_device(ctypes.byref(pb.desc), _ptr(pb.logits), _ptr(pb.targets), _ptr(pb.logits_lengths), _ptr(pb.targets_lengths), None, 0, 1.0, _ptr(grads), _ptr(ws), ws.numel(), _stream(xg.device))
torch.cuda.synchronize(); print("backward rc", rc, flush=True)
xr = x.clone().requires_grad_(); l = oracle.ctc_loss_module(oracle.engine(0), xr, tg, ll, tl, reduce=True); l.backward()
print("grad maxdiff", (grads.cpu() - xr.grad).abs().max().item(), flush=True)
