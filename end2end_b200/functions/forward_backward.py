"""Autograd bridge: the counterpart of pytorch_end2end/functions/forward_backward.py:4-35.

The reference Function calls ``engine.compute`` (a pybind11 C++ CPU engine), keeps the dense
gradient on ``ctx`` and multiplies it by ``grad_output`` in ``backward``.  This one calls the C ABI
(``end2end_b200.engine``): with CUDA tensors ``forward`` makes ONE library call that produces the
loss, its reduction and the gradient block with the constant part of ``grad_output`` (1, or 1/B for
a mean) folded in -- a single fused kernel for alphabets <= 128 -- and ``backward`` applies the
run-time ``grad_output`` in place with a kernel whose blocks return immediately when it is 1, so
the dense [B,T,V] block is written exactly once.  With CPU tensors the engine's
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
        ctx.inputs = None
        need_grad = ctx.needs_input_grad[1]
        scale = 1.0 / ctx.batch if reduction == "mean" else 1.0
        if logits.is_cuda:
            if need_grad:
                # one library call: loss, gradient (the constant part of grad_output already folded in)
                # and the reduction; backward only applies the run-time grad_output
                loss, grads, reduced, _ = engine.step(logits, targets, logits_lengths, targets_lengths, from_logits,
                                                      grad_scale=scale, reduce_scale=scale if reduction else None)
                ctx.grads, ctx.pristine = grads, True
                ctx.inputs = (logits, targets, logits_lengths, targets_lengths, from_logits, scale)
                if reduction is not None:
                    loss = reduced
            else:
                loss, _ = engine.forward(logits, targets, logits_lengths, targets_lengths, from_logits)
                ctx.grads = None
                if reduction is not None:
                    loss = engine.reduce(loss, scale)
        else:
            loss, grads = engine.compute(logits, targets, logits_lengths, targets_lengths, from_logits)
            ctx.grads, ctx.pristine = (grads if need_grad else None), False
            ctx.host_scale = scale
            if reduction is not None:
                loss = loss.mean() if reduction == "mean" else loss.sum()
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        """
        :param grad_output: [batch_size] (or 0-dim for a reduced loss)
        :return: gradient for logits, None for everything else
        """
        if ctx.inputs is not None:
            if not ctx.pristine:   # a second backward through a retained graph: rebuild the unscaled block
                logits, targets, ll, tl, from_logits, scale = ctx.inputs
                _, ctx.grads, _, _ = ctx.engine.step(logits, targets, ll, tl, from_logits, grad_scale=scale)
            ctx.pristine = False
            grad = ctx.engine.scale_rows_(ctx.grads, grad_output)
        else:
            g = grad_output.to(ctx.grads.device)
            g = g.reshape(-1, 1, 1) if g.numel() > 1 else g.reshape(1, 1, 1)
            grad = ctx.grads * (g * ctx.host_scale)
        return None, grad, None, None, None, None, None
