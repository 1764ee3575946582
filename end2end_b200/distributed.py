"""Batch-sharded CTC across the GPUs of one node (one process per GPU, torch.distributed / NCCL).

The reference has no multi-device path: its only parallelism is one host thread per utterance
(src/losses/forward_backward.cpp:36-53), and utterances share nothing.  So the path shards by
utterance with NO data-path collective: every GPU runs the full kernel pipeline on its own bucket
and owns those gradient rows.  The single exchange step is the reduction of
modules/ctc_loss.py:52-56 (``loss.sum()`` / ``loss.mean()``): one all-reduce of the pair
{local loss sum, local utterance count} (16 bytes) -- a pure-latency NCCL collective over
NVLink/NVSwitch, launched asynchronously so the gradient kernel does not wait for it when the
global batch size is known.

``plan_shards`` is the host-side planner: cost-balanced (longest-processing-time-first) buckets,
each bucket sorted by length so that CTAs scheduled together have similar trip counts.
"""
import torch
import torch.distributed as dist
import torch.nn as nn
from torch.autograd import Function

from .engine import CTCLossEngine


def utterance_cost(frames, target_len, alphabet):
    """Relative cost of one utterance: lattice cells (T * (2L+1)) plus the streamed row bytes."""
    return int(frames) * (2 * int(target_len) + 1) + int(frames) * int(alphabet) // 4


def plan_shards(logits_lengths, targets_lengths, alphabet, world_size):
    """LPT assignment of utterance indices to ``world_size`` buckets.

    Returns ``world_size`` lists of indices into the global batch; every index appears exactly
    once; within a bucket utterances are ordered by decreasing cost.  Deterministic.
    """
    ll = [int(v) for v in torch.as_tensor(logits_lengths).tolist()]
    tl = [int(v) for v in torch.as_tensor(targets_lengths).tolist()]
    if len(ll) != len(tl):
        raise ValueError("length vectors differ in size")
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    cost = [utterance_cost(a, b, alphabet) for a, b in zip(ll, tl)]
    order = sorted(range(len(cost)), key=lambda i: (-cost[i], i))
    buckets = [[] for _ in range(world_size)]
    load = [0] * world_size
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], len(buckets[k]), k))
        buckets[r].append(i)
        load[r] += cost[i]
    return buckets


def take_shard(indices, logits, targets, logits_lengths, targets_lengths, time_major=False):
    """Slice one rank's bucket out of a global batch (host-side helper for data loaders)."""
    idx = torch.as_tensor(indices, dtype=torch.int64, device=logits.device)
    lg = logits.index_select(1 if time_major else 0, idx)
    return (lg, targets.index_select(0, idx.to(targets.device)),
            logits_lengths.index_select(0, idx.to(logits_lengths.device)),
            targets_lengths.index_select(0, idx.to(targets_lengths.device)))


class LossComm:
    """The library-owned NCCL communicator for the one exchange of this path (``e2e_ctc_comm_*`` in
    include/e2e_ctc.h): an in-place all-reduce of the reduced loss scalar, enqueued on the caller's stream
    by the same library that enqueued the kernels.  Built once per process over the default process group
    (the 128-byte NCCL id travels through ``torch.distributed``); ``LossComm.get()`` returns ``None`` when
    there is nothing to build it from (no process group, a non-NCCL backend, libnccl missing), and the
    caller falls back to ``dist.all_reduce``."""
    _instance, _tried = None, False

    def __init__(self, handle, L):
        self.handle, self._L = handle, L

    @classmethod
    def get(cls):
        if cls._tried:
            return cls._instance
        cls._tried = True
        import ctypes
        import os
        from . import _lib
        if os.environ.get("E2E_CTC_NO_LIB_COMM") or not (dist.is_available() and dist.is_initialized()):
            return None
        if dist.get_backend() != "nccl" or dist.get_world_size() < 2 or not torch.cuda.is_available():
            return None
        # Measured anomaly (profiles/r01/scale_r01e.md): a 2-rank communicator created here on a box with MORE
        # than two GPUs (2 of 8 on an NVSwitch node) runs the scalar all-reduce 2-3x slower than c10d's own
        # communicator, while 2 of 2, 4 of 8 and 8 of 8 are faster than c10d.  Until that is understood the
        # partial-box two-rank case stays on c10d (E2E_CTC_LIB_COMM=1 forces the library communicator).  Cause found:
        # NVLS (NVLink SHARP multicast) set-up; with NCCL_NVLS_ENABLE=0 in the environment BEFORE the first NCCL
        # communicator of the process the two-rank library communicator is the fastest option again (0.160 ms/step).
        if (dist.get_world_size() == 2 and torch.cuda.device_count() > 2 and os.environ.get("NCCL_NVLS_ENABLE") != "0"
                and not os.environ.get("E2E_CTC_LIB_COMM")):
            return None
        L = _lib.load()
        rank, world = dist.get_rank(), dist.get_world_size()
        buf = ctypes.create_string_buffer(128)
        # EVERY rank checks that it can resolve libnccl before anyone enters the collective ncclCommInitRank: a rank
        # that returned early from e2e_ctc_comm_create would leave the others blocked inside it forever
        have = torch.tensor([1 if L.e2e_ctc_comm_unique_id(buf) == 0 else 0], device="cuda")
        dist.all_reduce(have, op=dist.ReduceOp.MIN)
        if int(have.item()) == 0:
            return None
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0)          # every rank takes rank 0's id
        h = ctypes.c_void_p()
        rc = L.e2e_ctc_comm_create(box[0], world, rank, ctypes.byref(h))
        flag = torch.tensor([1 if rc == 0 else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)     # all ranks or none
        if int(flag.item()) == 0:
            return None
        cls._instance = cls(h, L)
        return cls._instance

    def allreduce_sum_(self, t):
        """In-place sum of the CUDA tensor ``t`` over the ranks, on the current stream."""
        from . import _lib
        from .engine import _DTYPES, _stream
        _lib.check(self._L.e2e_ctc_comm_allreduce_sum(self.handle, t.data_ptr(), t.numel(), _DTYPES[t.dtype], _stream(t.device)))
        return t


class _ShardedLossFunction(Function):
    @staticmethod
    def forward(ctx, engine, logits, targets, logits_lengths, targets_lengths, from_logits, mean,
                group, global_batch):
        ctx.engine, ctx.mean = engine, mean
        need_grad = ctx.needs_input_grad[1]
        B = logits.size(0)
        known = float(global_batch) if global_batch is not None else None
        ctx.folded = 1.0 / known if (mean and known) else 1.0     # constant part of grad_output baked into the kernel
        ctx.on_device = logits.is_cuda
        ctx.redo, ctx.pristine = None, True
        reducing = group is not False and dist.is_available() and dist.is_initialized()
        if logits.is_cuda and need_grad and (known or not mean):
            # fast path (the global batch is known, or the loss is a plain sum): the reduce kernel already
            # leaves scale * sum(local losses) as a 0-dim tensor of the logits dtype; that scalar IS the
            # all-reduce buffer and the result -- no host-side arithmetic around the collective
            _, ctx.grads, total, _ = engine.step(logits, targets, logits_lengths, targets_lengths, from_logits,
                                                 grad_scale=ctx.folded, reduce_scale=ctx.folded)
            ctx.redo = (logits, targets, logits_lengths, targets_lengths, from_logits)
            if reducing:                                                    # THE collective of this path
                comm = LossComm.get() if group is None else None
                if comm is not None:
                    comm.allreduce_sum_(total)
                else:
                    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
            ctx.inv_n = ctx.folded if mean else None
            return total
        if logits.is_cuda and need_grad:
            _, ctx.grads, _, pair = engine.step(logits, targets, logits_lengths, targets_lengths, from_logits,
                                                grad_scale=ctx.folded, want_pair=True)
            ctx.redo = (logits, targets, logits_lengths, targets_lengths, from_logits)
        elif logits.is_cuda:
            losses, _ = engine.forward(logits, targets, logits_lengths, targets_lengths, from_logits)
            ctx.grads = None
            pair = engine.partial_sum(losses)           # fp64 {sum, B} written by the reduce kernel
        else:
            losses, grads = engine.compute(logits, targets, logits_lengths, targets_lengths, from_logits)
            ctx.grads = grads if need_grad else None
            pair = torch.stack([losses.double().sum(), torch.tensor(float(B), dtype=torch.float64)])
        if reducing:
            dist.all_reduce(pair, op=dist.ReduceOp.SUM, group=group)   # THE collective of this path
        ctx.inv_n = None
        if mean:
            ctx.inv_n = (1.0 / known) if known else (1.0 / pair[1])
            total = pair[0] * ctx.inv_n
        else:
            total = pair[0]
        return total.to(logits.dtype)

    @staticmethod
    def backward(ctx, grad_output):
        g = grad_output
        if ctx.mean and not isinstance(ctx.inv_n, float):
            g = g * ctx.inv_n.to(g.dtype)              # global count only known after the all-reduce
        if ctx.on_device:
            if not ctx.pristine and ctx.redo is not None:
                # a second backward through a retained graph: the block was scaled in place by the first one; rebuild it
                _, ctx.grads, _, _ = ctx.engine.step(*ctx.redo[:4], ctx.redo[4], grad_scale=ctx.folded)
            ctx.pristine = False
            grad = ctx.engine.scale_rows_(ctx.grads, g)
        else:
            grad = ctx.grads * (g.to(ctx.grads.device).reshape(1, 1, 1) * ctx.folded)
        return None, grad, None, None, None, None, None, None, None


class ShardedCTCLoss(nn.Module):
    """CTCLoss over a batch sharded across ranks: same arguments as ``CTCLoss`` plus the process
    group.  Each rank passes ITS bucket; the returned scalar is the global sum (``reduce=True``) or
    the global mean (``reduce=True, size_average=True``), identical on every rank; the gradient on
    each rank is the gradient of that global loss w.r.t. the rank's own logits.
    With ``reduce`` falsy the local per-utterance losses are returned and nothing is exchanged.

    :param global_batch: total utterances over all ranks, if known; lets the gradient kernel run
        without waiting for the all-reduce.
    """

    def __init__(self, size_average=None, reduce=None, after_logsoftmax=False, time_major=False, blank_idx=0,
                 process_group=None, global_batch=None, engine=None):
        super().__init__()
        self._reduce, self._size_average = reduce, size_average
        self._after_logsoftmax, self._time_major = after_logsoftmax, time_major
        self._group, self._global_batch = process_group, global_batch
        self._engine = engine if engine is not None else CTCLossEngine(blank_idx)

    def graphed(self, logits, targets, logits_lengths, targets_lengths, workspace=None):
        """The sharded step for device tensors at fixed addresses as CUDA graphs (SURVEY 8(f1) + 8(e)); needs
        ``reduce=True`` and, for a mean, ``global_batch``.  See :class:`ShardedGraphedStep`."""
        if not self._reduce:
            raise ValueError("graphed(): reduce=True required (nothing is exchanged otherwise: use CTCLoss.graphed)")
        mean = bool(self._size_average)
        if mean and self._global_batch is None:
            raise ValueError("graphed(): a mean over the ranks needs global_batch")
        if self._time_major:
            logits = logits.permute(1, 0, 2)
        scale = 1.0 / float(self._global_batch) if mean else 1.0
        return ShardedGraphedStep(self._engine, logits, targets, logits_lengths, targets_lengths,
                                  not self._after_logsoftmax, scale, self._group, self._time_major, workspace)

    def forward(self, logits, targets, logits_lengths, targets_lengths):
        if self._time_major:
            logits = logits.permute(1, 0, 2)
        if not self._reduce:
            from .functions.forward_backward import ForwardBackwardLossFunction
            return ForwardBackwardLossFunction.apply(self._engine, logits, targets, logits_lengths,
                                                     targets_lengths, not self._after_logsoftmax, None)
        return _ShardedLossFunction.apply(self._engine, logits, targets, logits_lengths, targets_lengths,
                                          not self._after_logsoftmax, bool(self._size_average),
                                          self._group, self._global_batch)


class ShardedGraphedStep:
    """The sharded training step of the loss, replayed as CUDA graphs, with the ONE collective of the path off the
    critical path: ``replay()`` launches the rank's captured step (memset, lattice kernel, reduction) on the current
    stream and enqueues the all-reduce of its scalar on a SIDE stream, so the collective of step k overlaps the
    lattice kernel of step k+1 -- the gradient never depends on it (the global batch size is folded in).  Two
    captured steps with their own scalar buffers alternate, so step k+1's reduction never overwrites the buffer
    step k's all-reduce is still reading; ``replay()`` makes the current stream wait only for the all-reduce of step
    k-2.  ``wait()`` joins the side stream; ``total`` then holds the global loss of the last replay."""

    _side_streams = {}     # ONE side stream per device for every step's all-reduce: NCCL operations of a communicator are
                   # issued in the same order on every rank, and that order is the replay order

    def __init__(self, engine, logits, targets, logits_lengths, targets_lengths, from_logits, scale, group, time_major,
                 workspace=None):
        a = engine.graphed_step(logits, targets, logits_lengths, targets_lengths, from_logits, grad_scale=scale, reduce_scale=scale,
                                workspace=workspace)
        b = engine.graphed_step(logits, targets, logits_lengths, targets_lengths, from_logits, grad_scale=scale, reduce_scale=scale, share=a)
        self._steps = (a, b)
        self._k = 0
        self._group = group
        self._reducing = group is not False and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self._comm = LossComm.get() if (self._reducing and group is None) else None
        key = logits.device.index if logits.device.index is not None else torch.cuda.current_device()
        if key not in ShardedGraphedStep._side_streams:
            ShardedGraphedStep._side_streams[key] = torch.cuda.Stream(device=logits.device)
        self._side = ShardedGraphedStep._side_streams[key]
        self._ev_step = [torch.cuda.Event(), torch.cuda.Event()]
        self._ev_red = [None, None]
        self.grad = a.grads.permute(1, 0, 2) if time_major else a.grads
        self.total = a.reduced

    def replay(self):
        i = self._k & 1
        self._k += 1
        step = self._steps[i]
        cur = torch.cuda.current_stream(step.device)
        if self._ev_red[i] is not None:
            cur.wait_event(self._ev_red[i])          # the all-reduce that last used this scalar buffer (two steps ago)
        step.launch()
        self.total = step.reduced
        if self._reducing:
            self._ev_step[i].record(cur)
            self._side.wait_event(self._ev_step[i])
            with torch.cuda.stream(self._side):
                if self._comm is not None:
                    self._comm.allreduce_sum_(step.reduced)
                else:
                    dist.all_reduce(step.reduced, op=dist.ReduceOp.SUM, group=self._group)
                ev = torch.cuda.Event()
                ev.record(self._side)
            self._ev_red[i] = ev
        return self.total, self.grad

    def wait(self):
        """Make the current stream wait for the outstanding all-reduces (call before reading ``total``)."""
        for ev in self._ev_red:
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)
        return self.total
