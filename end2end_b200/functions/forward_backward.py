"""Autograd bridge: the counterpart of pytorch_end2end/functions/forward_backward.py:4-35.

The reference Function calls ``engine.compute`` (a pybind11 C++ CPU engine), keeps the dense
gradient on ``ctx`` and multiplies it by ``grad_output`` in ``backward``.  This one calls the C ABI
(``end2end_b200.engine``): with CUDA tensors ``forward`` runs row statistics + the lattice and
keeps only the lattice workspace; ``backward`` runs the gradient kernel with ``grad_output``
folded in, so the dense [B,T,V] block is written exactly once.  With CPU tensors the engine's
host entry point produces loss and gradient in one round trip (as the reference's engine does)
and ``backward`` scales it.

``apply(engine, logits, targets, logits_lengths, targets_lengths[, from_logits[, reduction]])``
    from_logits: the input is raw logits and log_softmax is fused (CTCLoss(after_logsoftmax=False))
    reduction:   None -> losses [B]; "sum" / "mean" -> 0-dim reduced loss computed on the device
"""
import torch
from torch.autograd import Function


class ForwardBackwardLossFunction(Function):
    @staticmethod
    def forward(ctx, engine, logits, targets, logits_lengths, targets_lengths, from_logits=False,
                reduction=None):
        """
        :param engine: ``end2end_b200.engine.CTCLossEngine``
        :param logits: tensor [batch_size, sequence_length, alphabet_size] (any batch/time strides)
        :param targets: [batch_size, targets_sequence_length]
        :param logits_lengths: [batch_size]
        :param targets_lengths: [batch_size]
        :return: loss [batch_size], or a 0-dim tensor when ``reduction`` is "sum" / "mean"
        """
        ctx.engine = engine
        ctx.reduction = reduction
        ctx.batch = logits.size(0)
        need_grad = ctx.needs_input_grad[1]
        if logits.is_cuda:
            loss, state = engine.forward(logits, targets, logits_lengths, targets_lengths, from_logits)
            ctx.state = state if need_grad else None
            ctx.grads = None
            if reduction is not None:
                loss = engine.reduce(loss, 1.0 / ctx.batch if reduction == "mean" else 1.0)
        else:
            loss, grads = engine.compute(logits, targets, logits_lengths, targets_lengths, from_logits)
            ctx.state = None
            ctx.grads = grads if need_grad else None
            if reduction is not None:
                loss = loss.mean() if reduction == "mean" else loss.sum()
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        """
        :param grad_output: [batch_size] (or 0-dim for a reduced loss)
        :return: gradient for logits, None for everything else
        """
        scale = 1.0 / ctx.batch if ctx.reduction == "mean" else 1.0
        if ctx.state is not None:
            grad = ctx.engine.backward(ctx.state, grad_output, scale)
        else:
            g = grad_output.to(ctx.grads.device)
            g = g.reshape(-1, 1, 1) if g.numel() > 1 else g.reshape(1, 1, 1)
            grad = ctx.grads * (g * scale)
        return None, grad, None, None, None, None, None
