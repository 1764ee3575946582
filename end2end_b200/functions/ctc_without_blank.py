"""CTC without blank on the GPU behind the reference's Function
(pytorch_end2end/functions/ctc_without_blank.py:120-143).

The reference moves everything to numpy and runs one numba-jitted ``_ctc_without_blank_loss`` (:13-88) per
utterance on a Python thread (:91-117); here the batch is one kernel launch
(``e2e_ctc_noblank_fwd_bwd_device``, csrc/ctc_noblank.cu).  No CPU fallback.
"""
import ctypes

import torch
from torch.autograd import Function

from .. import _lib
from ..engine import _DTYPES, _ITYPES, _as_index, _dense3, _on_device, _ptr, _require_cuda, _stream


def ctc_without_blank_3d_loss(log_probs, targets, logits_lengths, targets_lengths, space_idx=-1):
    """(losses [B], grads [B,T,V]) on ``log_probs``' CUDA device, in its dtype: the device counterpart of the reference's
    ``_ctc_without_blank_3d_loss`` (:91-117).  ``grads`` is zero past every utterance's frames."""
    _require_cuda()
    L = _lib.load()
    lp = log_probs.detach()
    if lp.dim() != 3:
        raise ValueError("logits must be [batch, frames, alphabet]")
    if lp.dtype not in _DTYPES:
        raise TypeError("unsupported dtype %s" % lp.dtype)
    if not lp.is_cuda:
        lp = lp.cuda()
    dev = lp.device
    if not _dense3(lp):
        lp = lp.contiguous()
    B, T, V = lp.shape
    if targets.dim() != 2 or targets.size(0) != B:
        raise ValueError("targets must be [batch, max_target_length]")
    if not -1 <= space_idx < V:
        raise ValueError("space_idx %d outside [-1,%d)" % (space_idx, V))
    targets = _as_index(targets, dev, "targets")
    if targets.size(1) > 0 and targets.stride(1) != 1:
        targets = targets.contiguous()
    ll = _as_index(logits_lengths, dev, "logits_lengths").contiguous()
    tl = _as_index(targets_lengths, dev, "targets_lengths").contiguous()
    if ll.dtype != tl.dtype:
        ll, tl = ll.to(torch.int64), tl.to(torch.int64)
    d = _lib.Desc()
    d.batch, d.max_frames, d.alphabet, d.max_targets = B, T, V, targets.size(1)
    d.blank_idx, d.dtype = 0, _DTYPES[lp.dtype]
    d.targets_itype, d.lengths_itype = _ITYPES[targets.dtype], _ITYPES[ll.dtype]
    d.logits_stride_b, d.logits_stride_t = lp.stride(0), lp.stride(1)
    d.targets_stride_b = targets.stride(0) if targets.size(1) > 0 else 0
    with _on_device(dev):
        n = L.e2e_ctc_noblank_workspace_bytes(ctypes.byref(d))
        if n == 0:
            raise _lib.E2EError(1, L.e2e_last_error_string().decode("utf-8", "replace"))
        ws = torch.empty(n, dtype=torch.uint8, device=dev)
        losses = torch.empty(B, dtype=lp.dtype, device=dev)
        grads = torch.empty_strided(lp.size(), lp.stride(), dtype=lp.dtype, device=dev)
        d.grads_stride_b, d.grads_stride_t = grads.stride(0), grads.stride(1)
        _lib.check(L.e2e_ctc_noblank_fwd_bwd_device(ctypes.byref(d), int(space_idx), _ptr(lp), _ptr(targets), _ptr(ll), _ptr(tl),
                                                    _ptr(losses), _ptr(grads), _ptr(ws), ws.numel(), _stream(dev)))
    return losses, grads


class CTCWithoutBlankLossFunction(Function):
    """``apply(log_probs, targets, logits_lengths, targets_lengths, space_idx=-1) -> losses [B]`` (reference :120-143).
    The gradient block is produced in forward (as the reference does) and scaled by ``grad_output`` in backward."""

    @staticmethod
    def forward(ctx, log_probs, targets, logits_lengths, targets_lengths, space_idx=-1):
        losses, grads = ctc_without_blank_3d_loss(log_probs, targets, logits_lengths, targets_lengths, space_idx)
        on_host = not log_probs.is_cuda          # results live where the input lives
        ctx.grads = grads.cpu() if on_host else grads
        return losses.cpu() if on_host else losses

    @staticmethod
    def backward(ctx, grad_output):
        scale = grad_output.to(ctx.grads.device).reshape(-1, 1, 1)
        return ctx.grads * scale, None, None, None, None
