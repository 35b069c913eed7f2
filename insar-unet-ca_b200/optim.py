"""Fused multi-tensor Adam for the drop-in UNet: `optim.Adam(model.parameters(), lr)` (UCA:466)
and `optimizer.step()` (UCA:346).

    optimizer = unetca_b200.optim.Adam(model.parameters(), lr=LEARNING_RATE, model=model)

is a `torch.optim.Optimizer` with torch.optim.Adam's constructor, arithmetic and `state_dict` layout (`step`,
`exp_avg`, `exp_avg_sq` per parameter — checkpoints move freely between the two).  One step is a handful of launches
whatever the number of tensors (the reference model has 154): a one-thread kernel advances the step count and the
bias corrections on the device, a multi-tensor kernel steps the small tensors, and — when `model=` names the UNet the
parameters belong to — the 3x3 convolution filters (99 % of the parameters) are stepped by a kernel that also writes
their packed operand copies (forward `[O][tap*C+c]` and dgrad `[C][tap'*O+o]`, bf16 or fp32 by the model's precision)
straight into the model's operand cache, so the next forward repacks nothing.

The step count lives on the device (`hyper[0]`): the optimizer is capturable as is (`graph.GraphedTrainStep`).
All parameters of a group share one step count (every parameter of the reference gets a gradient every step); a
parameter whose `.grad` is None is skipped for that step, as in torch.

CUDA only: there is no CPU path.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
from torch.optim import Optimizer

from . import _lib


class Adam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False, *,
                 maximize=False, model=None):
        if amsgrad or maximize:
            raise NotImplementedError("unetca_b200.optim.Adam implements the reference's configuration: no amsgrad, no maximize")
        if isinstance(lr, torch.Tensor):
            raise NotImplementedError("unetca_b200.optim.Adam takes a float learning rate")
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError(f"invalid Adam hyper-parameters: lr={lr} betas={betas} eps={eps} weight_decay={weight_decay}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False,
                                      maximize=False))
        self._model = model
        self._hyper = {}               # group index -> 8 floats on the device (hyper[0] = step count)
        self._conv_of = {}             # id(param) -> nn.Conv2d whose filter the packing kernel handles
        if model is not None:
            for blk in list(model._enc_blocks) + list(model._dec_blocks):
                for conv in (blk.conv1, blk.conv2):
                    O, C = conv.weight.shape[0], conv.weight.shape[1]
                    if O % 32 == 0 and C % 32 == 0 and not (blk.first and conv is blk.conv1):
                        self._conv_of[id(conv.weight)] = conv

    # ---- state ---------------------------------------------------------------------------------------
    def _hyper_buf(self, gi, device):
        h = self._hyper.get(gi)
        if h is None or h.device != device:
            h = torch.zeros(8, dtype=torch.float32, device=device)
            self._hyper[gi] = h
        return h

    def _init_state(self, p):
        st = self.state[p]
        if "exp_avg" not in st:
            st["step"] = torch.zeros((), dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        return st

    def state_dict(self):
        """torch.optim.Adam's layout; the per-parameter `step` entries are read back from the device counters here."""
        for gi, group in enumerate(self.param_groups):
            h = self._hyper.get(gi)
            if h is None:
                continue
            t = float(h[0].item())
            for p in group["params"]:
                if p in self.state and "step" in self.state[p]:
                    self.state[p]["step"] = torch.tensor(t, dtype=torch.float32)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for gi, group in enumerate(self.param_groups):
            steps = [float(self.state[p]["step"]) for p in group["params"] if p in self.state and "step" in self.state[p]]
            if not steps:
                continue
            if max(steps) != min(steps):
                raise ValueError("unetca_b200.optim.Adam keeps one step count per parameter group; the loaded state has "
                                 f"{min(steps)}..{max(steps)}")
            dev = next(p.device for p in group["params"])
            self._hyper_buf(gi, dev)[0] = steps[0]
            for p in group["params"]:
                st = self.state.get(p)
                if st and "exp_avg" in st:
                    # own copies: torch's load_state_dict aliases tensors that already have the right dtype and device
                    st["exp_avg"] = st["exp_avg"].to(torch.float32).contiguous().clone()
                    st["exp_avg_sq"] = st["exp_avg_sq"].to(torch.float32).contiguous().clone()

    # ---- step ----------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        model = self._model
        eng = model._engine() if model is not None else None
        for gi, group in enumerate(self.param_groups):
            rows, conv_rows, adopted = [], [], []
            dev = None
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("unetca_b200.optim.Adam steps CUDA parameters only; there is no CPU fallback")
                if p.dtype != torch.float32 or g.dtype != torch.float32 or g.is_sparse:
                    raise RuntimeError("unetca_b200.optim.Adam needs dense fp32 parameters and gradients")
                if not p.is_contiguous():
                    raise RuntimeError("unetca_b200.optim.Adam needs contiguous parameters")
                if not g.is_contiguous():
                    g = g.contiguous()
                    p.grad = g
                dev = p.device
                st = self._init_state(p)
                conv = self._conv_of.get(id(p))
                if conv is not None:
                    dt, tdt = model._dt()
                    wf, wd, ldk, _, _ = eng.conv_w(conv, dt, tdt, False)        # the cached copies (built once)
                    conv_rows.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                      wf.data_ptr(), wd.data_ptr(), p.shape[0], p.shape[1]))
                    adopted.append(conv)
                else:
                    rows.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                 p.numel()))
            if dev is None:
                continue
            hyper = self._hyper_buf(gi, dev)
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream().cuda_stream
                b1, b2 = group["betas"]
                _lib.call("unetca_adam_tick", hyper.data_ptr(), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                          float(group["weight_decay"]), stream)
                if rows:
                    tab = np.asarray(rows, dtype=np.int64)
                    n = _lib.call("unetca_adam_step", tab.ctypes.data_as(ctypes.c_void_p), len(rows), hyper.data_ptr(), stream)
                    _lib.launch_count += n - 1
                if conv_rows:
                    dt, tdt = model._dt()
                    tab = np.asarray(conv_rows, dtype=np.int64)
                    _lib.call("unetca_adam_step_conv3x3", dt, tab.ctypes.data_as(ctypes.c_void_p), len(conv_rows),
                              hyper.data_ptr(), stream)
                    for conv in adopted:
                        eng.conv_w_adopt(conv, dt, tdt)
        return loss
