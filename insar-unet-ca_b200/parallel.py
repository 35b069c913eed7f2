"""Data-parallel training for the drop-in UNet: one process per GPU, bucketed gradient all-reduce overlapped with
backward (the reference has no multi-GPU code; this wraps `loss.backward()` of UCA:345 — SURVEY.md §8e).

Semantics are those of PyTorch DDP: the batch is split by rank, BatchNorm statistics stay per replica, parameter
gradients are averaged over ranks, initial parameters and buffers are broadcast from rank 0.

Mechanics.  Two flat fp32 buffers laid out in the order backward completes the gradients (`model.grad_order`), cut
into a few contiguous buckets:

  * `work`  — the backward kernels write every parameter gradient of the current backward straight into its slice;
  * `grads` — the accumulated gradients.  `p.grad` of every parameter IS a view of this buffer (set here, DDP's
    `gradient_as_bucket_view`), and the autograd Function hands autograd `None` for the parameters, so nothing is
    ever accumulated twice whatever `zero_grad(set_to_none=...)` the caller uses.

When the last gradient of a bucket has been enqueued, the bucket is folded into `grads` (`copy_` after a
`zero_grad(set_to_none=True)`, `add_` otherwise — torch's accumulate-into-.grad rule) and, on a synchronising
backward, `all_reduce(AVG, async_op=True)` of that `grads` slice is launched on the process group's own stream (NCCL
over NVLink/NVSwitch on the GPU box, gloo in the CPU tests), so the collective of bucket i runs under the backward
kernels of the layers above it; `finish()` waits for the handles.  The big buckets (down4 57 MB, conv1 28 MB) complete
mid-backward, so at most the last small bucket (inc) is exposed.

Gradient accumulation (BASELINE configs[2]: global batch 512 as 64-image micro-steps at every GPU count, so that the
train-mode BatchNorm batch is 64 everywhere): `set_accumulation(k)` makes every k-th backward the synchronising one,
`no_sync()` is DDP's context manager of the same name.  Non-synchronising backwards only add into `grads`.
"""
from __future__ import annotations

import contextlib

import torch
import torch.distributed as dist

from .model import UNet, grad_order

DEFAULT_BUCKET_MB = 25.0


def plan_buckets(names, numels, bucket_mb=DEFAULT_BUCKET_MB):
    """Cut the completion-ordered gradient list into contiguous buckets of >= bucket_mb MB (the last may be smaller).
    Returns (offsets {name: (start, numel)}, bucket_bounds [(start, end, last_name)], total)."""
    limit = int(bucket_mb * 1024 * 1024 / 4)
    offsets, bounds = {}, []
    pos = start = 0
    for n, k in zip(names, numels):
        offsets[n] = (pos, k)
        pos += k
        if pos - start >= limit:
            bounds.append((start, pos, n))
            start = pos
    if pos > start:
        bounds.append((start, pos, names[-1]))
    return offsets, bounds, pos


class GradBuckets:
    """Gradient sink factory for `UNet` (see model._GradSink): flat buckets + overlapped all-reduce."""

    def __init__(self, model: UNet, process_group=None, bucket_mb: float = DEFAULT_BUCKET_MB, broadcast: bool = True):
        self.model = model
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.names = grad_order(model)
        self.params = dict(model.named_parameters())
        assert sorted(self.names) == sorted(self.params), "grad_order does not cover the model's parameters"
        self.shapes = {n: self.params[n].shape for n in self.names}
        self.offsets, self.bounds, total = plan_buckets(self.names, [self.params[n].numel() for n in self.names], bucket_mb)
        dev = next(model.parameters()).device
        self.work = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(total, dtype=torch.float32, device=dev)
        self._closing = {last: i for i, (_, _, last) in enumerate(self.bounds)}
        self._members = []                       # bucket -> parameter names
        for s, e, _ in self.bounds:
            self._members.append([n for n in self.names if s <= self.offsets[n][0] < e])
        self.avg = dist.ReduceOp.AVG if dev.type == "cuda" else dist.ReduceOp.SUM     # gloo has no AVG
        self.accumulation = 1
        self._micro = 0
        self._sync_override = None
        if broadcast and self.world > 1:
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)
            if hasattr(model, "invalidate_packed"):
                model.invalidate_packed()        # the broadcast wrote through .data: no version bump to see it by
        object.__setattr__(model, "_grad_sink_factory", self._new_sink)

    # ---- accumulation control -------------------------------------------------------------------------
    def set_accumulation(self, k: int) -> "GradBuckets":
        """Every k-th backward all-reduces; the k-1 before it only accumulate locally (micro-steps)."""
        if k < 1:
            raise ValueError("accumulation steps must be >= 1")
        self.accumulation, self._micro = int(k), 0
        return self

    @contextlib.contextmanager
    def no_sync(self):
        """Backwards inside only accumulate locally (torch DDP's no_sync)."""
        prev, self._sync_override = self._sync_override, False
        try:
            yield
        finally:
            self._sync_override = prev

    def _will_sync(self) -> bool:
        if self._sync_override is not None:
            return self._sync_override
        self._micro += 1
        if self._micro >= self.accumulation:
            self._micro = 0
            return True
        return False

    def view(self, name):
        s, k = self.offsets[name]
        return self.grads[s:s + k].view(self.shapes[name])

    def detach(self):
        object.__setattr__(self.model, "_grad_sink_factory", None)

    def _new_sink(self):
        return _BucketSink(self, self._will_sync())


class _BucketSink:
    owns_grads = True          # p.grad is bound here; the autograd Function returns None for the parameters

    def __init__(self, owner: GradBuckets, sync: bool):
        self.o = owner
        self.sync = sync
        self.handles = []
        self.next = 0
        # torch's rule for .grad: None -> becomes the new gradient, else accumulate.  Decided per bucket up front.
        self.fresh = []
        for members in owner._members:
            state = []
            for n in members:
                p, v = owner.params[n], owner.view(n)
                g = p.grad
                if g is None:
                    state.append(True)
                else:
                    if g.data_ptr() != v.data_ptr():          # a foreign .grad (set by the caller): adopt its values
                        v.copy_(g)
                    state.append(False)
            if any(state) and not all(state):
                for n, st in zip(members, state):
                    if st:
                        owner.view(n).zero_()
            self.fresh.append(all(state))

    def alloc(self, name, like):
        start, k = self.o.offsets[name]
        return self.o.work[start:start + k].view(self.o.shapes[name])

    def put(self, name):
        o = self.o
        assert o.names[self.next] == name, f"gradient completion order changed: expected {o.names[self.next]}, got {name}"
        self.next += 1
        b = o._closing.get(name)
        if b is None:
            return
        s, e, _ = o.bounds[b]
        if self.fresh[b]:
            o.grads[s:e].copy_(o.work[s:e])
        else:
            o.grads[s:e].add_(o.work[s:e])
        if self.sync and o.world > 1:
            self.handles.append(dist.all_reduce(o.grads[s:e], op=o.avg, group=o.pg, async_op=True))

    def finish(self):
        o = self.o
        assert self.next == len(o.names)
        for h in self.handles:
            h.wait()
        if self.sync and o.world > 1 and o.avg == dist.ReduceOp.SUM:
            o.grads.div_(o.world)
        for n, p in o.params.items():
            v = o.view(n)
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v
        return None


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rank's contiguous slice of a global batch (B must divide by world)."""
    B = x.shape[0]
    if B % world:
        raise ValueError(f"global batch {B} does not divide over {world} ranks")
    per = B // world
    return x[rank * per:(rank + 1) * per]
