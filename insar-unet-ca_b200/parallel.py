"""Data-parallel training for the drop-in UNet: one process per GPU, bucketed gradient all-reduce overlapped with
backward (the reference has no multi-GPU code; this wraps `loss.backward()` of UCA:345 — SURVEY.md §8e).

Semantics are those of PyTorch DDP: the batch is split by rank, BatchNorm statistics stay per replica, parameter
gradients are averaged over ranks, initial parameters and buffers are broadcast from rank 0.

Mechanics: every parameter gradient is written by the backward kernels *directly* into a slice of one flat fp32
buffer laid out in the order backward completes them (`model.grad_order`), cut into a few contiguous buckets.
When the last gradient of a bucket has been enqueued, `all_reduce(AVG, async_op=True)` is launched on the process
group's own stream (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests), so the collective of bucket i
runs under the backward kernels of the layers above it; `finish()` waits for the handles.  The big buckets
(down4 57 MB, conv1 28 MB) complete mid-backward, so at most the last small bucket (inc) is exposed.
"""
from __future__ import annotations

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
        self.world = dist.get_world_size(process_group)
        self.names = grad_order(model)
        params = dict(model.named_parameters())
        assert sorted(self.names) == sorted(params), "grad_order does not cover the model's parameters"
        self.shapes = {n: params[n].shape for n in self.names}
        self.offsets, self.bounds, total = plan_buckets(self.names, [params[n].numel() for n in self.names], bucket_mb)
        dev = next(model.parameters()).device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self._closing = {last: i for i, (_, _, last) in enumerate(self.bounds)}
        self.avg = dist.ReduceOp.AVG if dev.type == "cuda" else dist.ReduceOp.SUM     # gloo has no AVG
        if broadcast:
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)
        object.__setattr__(model, "_grad_sink_factory", self._new_sink)

    def detach(self):
        object.__setattr__(self.model, "_grad_sink_factory", None)

    def _new_sink(self):
        return _BucketSink(self)


class _BucketSink:
    def __init__(self, owner: GradBuckets):
        self.o = owner
        self.handles = []
        self.next = 0

    def alloc(self, name, like):
        start, k = self.o.offsets[name]
        return self.o.flat[start:start + k].view(self.o.shapes[name])

    def put(self, name):
        o = self.o
        assert o.names[self.next] == name, f"gradient completion order changed: expected {o.names[self.next]}, got {name}"
        self.next += 1
        b = o._closing.get(name)
        if b is not None and o.world > 1:
            s, e, _ = o.bounds[b]
            self.handles.append(dist.all_reduce(o.flat[s:e], op=o.avg, group=o.pg, async_op=True))

    def finish(self):
        o = self.o
        assert self.next == len(o.names)
        for h in self.handles:
            h.wait()
        if o.world > 1 and o.avg == dist.ReduceOp.SUM:
            o.flat.div_(o.world)
        return {n: o.flat[s:s + k].view(o.shapes[n]) for n, (s, k) in o.offsets.items()}


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rank's contiguous slice of a global batch (B must divide by world)."""
    B = x.shape[0]
    if B % world:
        raise ValueError(f"global batch {B} does not divide over {world} ranks")
    per = B // world
    return x[rank * per:(rank + 1) * per]
