"""Experiment: the fused kernel reading / writing PINNED HOST memory directly (UVA) instead of staged copies.
    python tools/zero_copy_probe.py [cfg]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle
from end2end_b200 import CTCLossEngine, _lib
from end2end_b200.engine import _Problem, _ptr, _stream
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
B, T, V, Lmin, Lmax, seed, dtype, full = oracle.CONFIGS[cfg]
x, tg, ll, tl = oracle.make_inputs(B, T, V, Lmin, Lmax, seed, dtype=dtype, full_length=full)
eng = CTCLossEngine(0)
L = _lib.load()
dev = torch.device("cuda", 0)
xp, tgp, llp, tlp = x.pin_memory(), tg.pin_memory(), ll.pin_memory(), tl.pin_memory()
xd = torch.empty_like(x, device=dev); tgd, lld, tld = tg.to(dev), ll.to(dev), tl.to(dev)
gd = torch.empty_like(xd); ld = torch.empty(B, dtype=dtype, device=dev)
gh = torch.empty_like(x).pin_memory(); lh = torch.empty(B, dtype=dtype).pin_memory()
pb = _Problem(0, xd, tgd, lld, tld, True, dev)
ws = torch.empty(eng._ws_need(pb), dtype=torch.uint8, device=dev)
st = _stream(dev)


def call(logits, targets, il, tl_, losses, grads):
    rc = L.e2e_ctc_loss_fwd_bwd_device(ctypes.byref(pb.desc), _ptr(logits), _ptr(targets), _ptr(il), _ptr(tl_), _ptr(losses), _ptr(grads), _ptr(ws), ws.numel(), st)
    assert rc == 0, L.e2e_last_error_string()


def timeit(fn, n=60):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def staged():
    xd.copy_(xp, non_blocking=True); tgd.copy_(tgp, non_blocking=True); lld.copy_(llp, non_blocking=True); tld.copy_(tlp, non_blocking=True)
    call(xd, tgd, lld, tld, ld, gd)
    gh.copy_(gd, non_blocking=True); lh.copy_(ld, non_blocking=True)
    torch.cuda.synchronize()


def out_zero_copy():
    xd.copy_(xp, non_blocking=True); tgd.copy_(tgp, non_blocking=True); lld.copy_(llp, non_blocking=True); tld.copy_(tlp, non_blocking=True)
    call(xd, tgd, lld, tld, lh, gh)
    torch.cuda.synchronize()


def all_zero_copy():
    call(xp, tgp, llp, tlp, lh, gh)
    torch.cuda.synchronize()


def in_zero_copy_small():      # logits staged, small tensors and outputs zero-copy
    xd.copy_(xp, non_blocking=True)
    call(xd, tgp, llp, tlp, lh, gh)
    torch.cuda.synchronize()


staged(); ref_g, ref_l = gh.clone(), lh.clone()
for name, fn in (("staged copies (torch)", staged), ("outputs zero-copy", out_zero_copy), ("logits staged, everything else zero-copy", in_zero_copy_small), ("everything zero-copy", all_zero_copy)):
    gh.zero_(); lh.zero_()
    fn()
    same = torch.equal(gh, ref_g) and torch.equal(lh, ref_l)
    t = timeit(fn)
    print("%-45s %.1f us/call (%.0f utt/s) results identical: %s" % (name, t, B / t * 1e6, same))
print("engine.compute (pinned inputs): %.1f us" % timeit(lambda: eng.compute(xp, tgp, llp, tlp, from_logits=True)))
