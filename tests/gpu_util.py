"""Helpers for the -m gpu parity tests: move NCHW CPU tensors to NHWC device buffers and call the C ABI."""
import ctypes

import torch

import unetca_b200
from unetca_b200 import _lib

F32, BF16 = _lib.F32, _lib.BF16
TDT = {F32: torch.float32, BF16: torch.bfloat16}


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def to_nhwc(x, dt):
    """(B,C,H,W) CPU float -> contiguous (B,H,W,C) CUDA tensor of storage type dt."""
    return x.permute(0, 2, 3, 1).contiguous().to(TDT[dt]).cuda()


def from_nhwc(t):
    return t.float().cpu().permute(0, 3, 1, 2).contiguous()


def rounded(x, dt):
    """x as it reads back after being stored in dt."""
    return x.to(TDT[dt]).float()


def parts_buf(B, width=2048):
    return torch.empty(_lib.load().unetca_max_parts(B) * width, dtype=torch.float32, device="cuda")


def call(name, *a):
    return _lib.call(name, *a)


def cint():
    return ctypes.c_int(0)


def relerr(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
