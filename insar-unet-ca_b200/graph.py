"""Whole-train-step CUDA graph (forward + loss + backward + optimizer step as one replayable launch).

At the reference's own configuration (batch 8, 1x128x128 tiles, UCA:21-24) a step is ~0.3 ms of GPU work behind ~280
kernel launches issued from Python: the host, not the GPU, sets the step time.  Every launch of this path goes to
PyTorch's current stream through the C ABI with caller-owned memory and no host synchronisation, so the step can be
captured once with `torch.cuda.graph` and replayed: the tensor maps (passed by value as kernel parameters), the scratch
buffers and the gradients all live at fixed addresses inside the graph's private memory pool.

    step = GraphedTrainStep(model, optimizer, images, masks)     # optimizer must be capturable (Adam(capturable=True))
    for images, masks in loader:
        loss = step(images, masks)                                # copies into the static inputs, replays the graph

Same kernels in the same order as the eager step, hence bit-identical results (tests/test_gpu_model.py), also when
eval-mode forwards (validation) are interleaved with replays.  Single GPU
only: the data-parallel bucket all-reduce is not captured.  If eager steps ran before, drop every reference to their
losses / outputs first (`loss = None`): autograd keeps the parameters' AccumulateGrad nodes bound to the stream they
were created on while an old graph is alive, and a capture that touches them is invalidated.
"""
from __future__ import annotations

import torch


class GraphedTrainStep:
    def __init__(self, model, optimizer, images: torch.Tensor, masks: torch.Tensor, ignore_index: int = 255, warmup: int = 3):
        if not images.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors; there is no CPU fallback")
        if model._grad_sink_factory is not None:
            raise NotImplementedError("GraphedTrainStep does not capture the data-parallel gradient all-reduce")
        self.model, self.optimizer, self.ignore_index = model, optimizer, ignore_index
        self.images, self.masks = images.clone(), masks.clone()
        model.train()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up off the default stream, as torch.cuda.graph requires
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self._eager(zero=False)
        # The captured kernels read and write the model's packed operand copies and scratch buffers by address.  The engine
        # rewrites those buffers in place (it never reallocates them while the parameters stay where they are); holding
        # them here as well keeps the memory alive even if the engine's caches are dropped.
        eng = model._engine()
        self._pinned = [list(eng._packed.values()), list(eng._scratch.values())]

    def _eager(self, zero=True):
        if zero:
            self.optimizer.zero_grad(set_to_none=True)
        loss = self.model.loss(self.images, self.masks, self.ignore_index)
        loss.backward()
        self.optimizer.step()
        return loss

    def __call__(self, images: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        self.images.copy_(images, non_blocking=True)
        self.masks.copy_(masks, non_blocking=True)
        self.graph.replay()
        # A replay steps the parameters without running any Python: neither Tensor._version nor the engine's own step
        # marker moves.  Tell the engine a train step happened, so the next eval-mode forward (the reference's per-epoch
        # validate_model, UCA:273-287) re-derives the packed filters from the stepped parameters instead of reusing the
        # ones of the previous validation.
        self.model._engine().last_train = True
        return self.loss
