"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the CTC hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the reported CPU
baseline.  The product (``end2end_b200`` / ``pytorch_end2end``) never imports it: there is no CPU
fallback on the product path.

Two engines are exposed behind one interface (``.compute(logits, targets, logits_lengths,
targets_lengths) -> (losses, grads)``, the signature of the reference's pybind11 class
``cpp_ctc_loss.CTCLossEngine``, src/losses/ctc_loss_py.cpp:5-17):

* ``PortEngine``  -- the plain-C restatement in ``oracle/ctc_oracle.c`` (kind ``"port"``),
* ``ref_engine()`` -- the UNMODIFIED reference C++ compiled by ``oracle/build_ref.py`` into
  ``oracle/_ref/`` (kind ``"reference"``); present when that directory travelled with the repo.

``beam_decode`` is the LM-free prefix beam search (decoders/ctc_decoder.py:76-115 over ctc_decoder.cpp:153-198,353-441)
through the compiled reference or the C restatement.  ``ctc_loss_module`` / ``greedy_decode`` restate the few Python lines the reference wraps around
its engines (pytorch_end2end/modules/ctc_loss.py:25-57, functions/forward_backward.py:6-35,
decoders/ctc_decoder.py:117-149) so either engine can be driven exactly as the reference drives it.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "ctc_oracle.c")
_LIB = os.path.join(_HERE, "libctc_oracle.so")
_lib = None


def build(force=False):
    """gcc the C restatement into oracle/libctc_oracle.so (a checker, never shipped as product)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-o", _LIB, _SRC, "-lm"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)
        L.ctc_oracle_batch.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, ctypes.c_int,
                                       ip, ip, ctypes.c_int, dp, dp]
        L.ctc_oracle_batch.restype = ctypes.c_int
        L.ctc_oracle_greedy.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, ctypes.c_int, ip, ip]
        L.ctc_oracle_greedy.restype = None
        L.ctc_oracle_log_softmax_f32.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_int64,
                                                 ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
        L.ctc_oracle_log_softmax_f32.restype = None
        L.ctc_oracle_align.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, ctypes.c_int, ip, ip,
                                       ctypes.c_int, ctypes.c_int, ip]
        L.ctc_oracle_align.restype = None
        L.ctc_oracle_noblank.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, ctypes.c_int,
                                         ip, ip, ctypes.c_int, dp, dp]
        L.ctc_oracle_noblank.restype = None
        L.ctc_oracle_beam.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_double, ctypes.c_double, ip, ip, ip]
        L.ctc_oracle_beam.restype = None
        _lib = L
    return _lib


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _iptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


class PortEngine:
    """C-oracle stand-in for ``cpp_ctc_loss.CTCLossEngine`` (forward_backward.cpp:7-59): widen to
    float64 on the CPU, run every utterance, cast loss and grads back to the caller's dtype."""

    kind = "port"

    def __init__(self, blank_idx):
        self.blank_idx = int(blank_idx)

    def compute(self, logits, targets, logits_lengths, targets_lengths):
        src_dtype, src_dev = logits.dtype, logits.device
        lp = np.ascontiguousarray(logits.detach().to("cpu").to(torch.float64).numpy())
        tg = np.ascontiguousarray(targets.to("cpu").to(torch.int64).numpy())
        il = np.ascontiguousarray(logits_lengths.to("cpu").to(torch.int64).numpy())
        tl = np.ascontiguousarray(targets_lengths.to("cpu").to(torch.int64).numpy())
        B, T, V = lp.shape
        if tg.ndim != 2:
            tg = tg.reshape(B, -1)
        Lmax = tg.shape[1]
        if Lmax == 0:
            tg = np.zeros((B, 1), dtype=np.int64)
            Lmax = 1
        losses = np.zeros(B, dtype=np.float64)
        grads = np.zeros_like(lp)
        rc = lib().ctc_oracle_batch(_dptr(lp), B, T, V, _iptr(tg), Lmax, _iptr(il), _iptr(tl),
                                    self.blank_idx, _dptr(losses), _dptr(grads))
        if rc != 0:
            raise ValueError("oracle: lengths/labels out of range (undefined behaviour in the reference)")
        return (torch.from_numpy(losses).to(src_dev).to(src_dtype),
                torch.from_numpy(grads).to(src_dev).to(src_dtype))


_ref_mods = {}


def _load_ref(name, sub):
    if name in _ref_mods:
        return _ref_mods[name]
    path = os.path.join(_HERE, "_ref", sub, name + ".so")
    mod = None
    if os.path.exists(path):
        import importlib.util
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules.setdefault(name, mod)
    _ref_mods[name] = mod
    return mod


def have_ref():
    return _load_ref("cpp_ctc_loss", "loss") is not None


def ref_engine(blank_idx):
    """The unmodified reference ``CTCLossEngine`` (compiled from /root/reference by build_ref.py)."""
    mod = _load_ref("cpp_ctc_loss", "loss")
    if mod is None:
        raise RuntimeError("oracle/_ref/loss/cpp_ctc_loss.so not built (run oracle/build_ref.py)")
    eng = mod.CTCLossEngine(int(blank_idx))
    eng_kind = "reference"
    return _Tagged(eng, eng_kind)


class _Tagged:
    def __init__(self, eng, kind):
        self._eng, self.kind = eng, kind

    def compute(self, *a):
        return self._eng.compute(*a)


def engine(blank_idx, prefer="reference"):
    if prefer == "reference" and have_ref():
        return ref_engine(blank_idx)
    return PortEngine(blank_idx)


class _FwdBwd(torch.autograd.Function):
    """functions/forward_backward.py:6-35: grads are produced in forward, scaled in backward."""

    @staticmethod
    def forward(ctx, eng, lp, targets, ll, tl):
        loss, grads = eng.compute(lp, targets, ll, tl)
        ctx.grads = grads
        return loss

    @staticmethod
    def backward(ctx, g):
        return None, ctx.grads.contiguous() * g.contiguous().view(-1, 1, 1), None, None, None


def ctc_loss_module(eng, logits, targets, logits_lengths, targets_lengths, reduce=None,
                    size_average=None, after_logsoftmax=False, time_major=False):
    """modules/ctc_loss.py:25-57 around an oracle engine (autograd flows to ``logits``)."""
    lp = logits if after_logsoftmax else F.log_softmax(logits, dim=2)
    if time_major:
        lp = lp.permute(1, 0, 2)
    loss = _FwdBwd.apply(eng, lp, targets, logits_lengths, targets_lengths)
    if reduce:
        return loss.mean() if size_average else loss.sum()
    return loss


def greedy_decode(logits, logits_lengths=None, blank_idx=0, time_major=False, labels=None, prefer="reference"):
    """decoders/ctc_decoder.py:117-149 + ctc_decoder.cpp:443-490.  Returns (targets[B,T] i64,
    lengths[B] i64, sentences)."""
    if time_major:
        logits = logits.transpose(1, 0)
    logits = logits.detach().cpu()
    B, T, V = logits.shape
    if logits_lengths is None:
        logits_lengths = torch.zeros(B, dtype=torch.int).fill_(T)
    logits_lengths = logits_lengths.cpu()
    labels = list(labels or [])
    mod = _load_ref("cpp_ctc_decoder", "decoder") if prefer == "reference" else None
    if mod is not None:
        dec = mod.CTCDecoder(int(blank_idx), 1, labels, "", 1.0, 0.0, -1000.0, False)
        return dec.decode_greedy(logits_=logits.contiguous(), logits_lengths_=logits_lengths)
    x = np.ascontiguousarray(logits.to(torch.float64).numpy())
    il = np.ascontiguousarray(logits_lengths.to(torch.int64).numpy())
    out = np.zeros((B, T), dtype=np.int64)
    out_len = np.zeros(B, dtype=np.int64)
    lib().ctc_oracle_greedy(_dptr(x), B, T, V, _iptr(il), int(blank_idx), _iptr(out), _iptr(out_len))
    sents = ["".join(labels[i] for i in out[b, :out_len[b]]) if labels else "" for b in range(B)]
    return torch.from_numpy(out), torch.from_numpy(out_len), sents


def log_softmax_f32(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    V = x.shape[-1]
    fp = ctypes.POINTER(ctypes.c_float)
    lib().ctc_oracle_log_softmax_f32(x.ctypes.data_as(fp), x.size // V, V, out.ctypes.data_as(fp))
    return out


def make_inputs(B, T, V, Lmin, Lmax, seed, dtype=torch.float32, full_length=False, scale=1.0):
    """The seeded synthetic draw of SURVEY.md section 8(d): randn logits, then target lengths, then
    targets in [1,V) (blank 0 never a target), then frame lengths in [3T/4, T]."""
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(B, T, V, generator=g) * scale).to(dtype)
    tl = torch.randint(Lmin, Lmax + 1, (B,), generator=g)
    tg = torch.randint(1, V, (B, Lmax), generator=g)
    if full_length:
        ll = torch.full((B,), T, dtype=torch.int64)
    else:
        ll = torch.randint(3 * T // 4, T + 1, (B,), generator=g)
    return x, tg, ll, tl


CONFIGS = {
    # name: (B, T, V, Lmin, Lmax, seed, dtype, full_length)   -- BASELINE.json configs[0..4]
    "c1": (4, 50, 28, 10, 29, 0, torch.float32, True),
    "c2": (64, 400, 29, 100, 200, 1, torch.float32, False),
    "c3": (1024, 128, 96, 20, 40, 2, torch.bfloat16, False),
    "c4": (128, 250, 1024, 40, 80, 3, torch.float32, False),
    "c5": (2048, 1600, 29, 300, 600, 4, torch.float32, False),
}


def get_alignment_3d(log_probs, targets, logits_lengths, targets_lengths, is_ctc=True, blank_idx=0):
    """pytorch_end2end/utils/alignment.py:109-138 through the C restatement: CPU int64 [B, T], -100 past the frames."""
    lp = np.ascontiguousarray(log_probs.detach().to("cpu").to(torch.float64).numpy())
    B, T, V = lp.shape
    tg = np.ascontiguousarray(targets.to("cpu").to(torch.int64).numpy()).reshape(B, -1)
    Lmax = tg.shape[1]
    if Lmax == 0:
        tg, Lmax = np.zeros((B, 1), dtype=np.int64), 1
    il = np.ascontiguousarray(logits_lengths.to("cpu").to(torch.int64).numpy())
    tl = np.ascontiguousarray(targets_lengths.to("cpu").to(torch.int64).numpy())
    out = np.zeros((B, T), dtype=np.int64)
    lib().ctc_oracle_align(_dptr(lp), B, T, V, _iptr(tg), Lmax, _iptr(il), _iptr(tl), int(blank_idx), 1 if is_ctc else 0, _iptr(out))
    return torch.from_numpy(out)


def ctc_without_blank(log_probs, targets, logits_lengths, targets_lengths, space_idx=-1):
    """pytorch_end2end/functions/ctc_without_blank.py:91-117 through the C restatement: (losses [B] float64,
    grads [B,T,V] float64; the reference casts both to float32)."""
    lp = np.ascontiguousarray(log_probs.detach().to("cpu").to(torch.float32).numpy())
    B, T, V = lp.shape
    tg = np.ascontiguousarray(targets.to("cpu").to(torch.int64).numpy()).reshape(B, -1)
    Lmax = tg.shape[1]
    if Lmax == 0:
        tg, Lmax = np.zeros((B, 1), dtype=np.int64), 1
    il = np.ascontiguousarray(logits_lengths.to("cpu").to(torch.int64).numpy())
    tl = np.ascontiguousarray(targets_lengths.to("cpu").to(torch.int64).numpy())
    losses = np.zeros(B, dtype=np.float64)
    grads = np.zeros((B, T, V), dtype=np.float64)
    lib().ctc_oracle_noblank(lp.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), B, T, V, _iptr(tg), Lmax, _iptr(il), _iptr(tl),
                             int(space_idx), _dptr(losses), _dptr(grads))
    return torch.from_numpy(losses), torch.from_numpy(grads)


def _space_id(labels):
    # src/decoders/ctc_decoder.cpp:56-60: index of " " in labels, -1 without labels / without a space
    labels = list(labels or [])
    return labels.index(" ") if " " in labels else -1


def beam_decode(logits, logits_lengths=None, blank_idx=0, beam_width=100, time_major=False, labels=None,
                after_logsoftmax=False, wip=0.0, oov_penalty=-1000.0, prefer="reference", return_ties=False):
    """decoders/ctc_decoder.py:76-115 (decode, LM-free) + ctc_decoder.cpp:153-198,353-441.  Returns (targets
    [B, max length] i64 zero padded, lengths [B] i64, sentences).  ``prefer="reference"`` runs the compiled
    reference when oracle/_ref travelled with the repo, else the C restatement."""
    with torch.no_grad():
        if not after_logsoftmax:
            logits = F.log_softmax(logits, -1)
        if time_major:
            logits = logits.transpose(1, 0)
    logits = logits.detach().cpu()
    B, T, V = logits.shape
    if logits_lengths is None:
        logits_lengths = torch.zeros(B, dtype=torch.int).fill_(T)
    logits_lengths = logits_lengths.cpu()
    labels = list(labels or [])
    mod = _load_ref("cpp_ctc_decoder", "decoder") if prefer == "reference" else None
    if mod is not None:
        dec = mod.CTCDecoder(int(blank_idx), int(beam_width), labels, "", 1.0, float(wip), float(oov_penalty), False)
        return dec.decode(logits_=logits.contiguous(), logits_lengths_=logits_lengths)
    x = np.ascontiguousarray(logits.to(torch.float64).numpy())
    il = np.ascontiguousarray(logits_lengths.to(torch.int64).numpy())
    Tw = max(T, 1)
    out = np.zeros((B, Tw), dtype=np.int64)
    out_len = np.zeros(B, dtype=np.int64)
    ties = np.zeros(B, dtype=np.int64)
    lib().ctc_oracle_beam(_dptr(x), B, T, V, _iptr(il), int(blank_idx), int(beam_width), _space_id(labels),
                          float(wip), float(oov_penalty), _iptr(out), _iptr(out_len), _iptr(ties))
    out = out[:, :max(int(out_len.max()), 0)] if B else out
    sents = ["".join(labels[i] for i in out[b, :out_len[b]] if i >= 0) if labels else "" for b in range(B)]
    if return_ties:
        return torch.from_numpy(np.ascontiguousarray(out)), torch.from_numpy(out_len), sents, torch.from_numpy(ties)
    return torch.from_numpy(np.ascontiguousarray(out)), torch.from_numpy(out_len), sents
