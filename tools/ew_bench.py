"""Micro-benchmark of the HBM-bound kernels on the U-Net-CA layer shapes (C ABI, CUDA events) under the tuning knobs.
usage: python tools/ew_bench.py [--B 64] [--ew-px 16 32 64] [--red-waves 0 1 2] [--shapes C,S ...]"""
import argparse
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unetca_b200 import _lib  # noqa: E402

SHAPES = [(64, 512), (128, 256), (256, 128), (512, 64), (1024, 32)]


def timeit(fn, iters):
    for _ in range(2):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--ew-px", type=int, nargs="*", default=[16])
    ap.add_argument("--red-waves", type=int, nargs="*", default=[1])
    ap.add_argument("--pool-quads", type=int, nargs="*", default=[4])
    ap.add_argument("--shapes", nargs="*", default=None)
    ap.add_argument("--only", default=None, help="substring filter on the kernel label")
    ap.add_argument("--apply-stream", type=int, default=None, help="tuning key 3: 8 KB tiles per block of the streamed bn_bwd_apply (0 = register kernel)")
    a = ap.parse_args()
    lib = _lib.load()
    if a.apply_stream is not None:
        lib.unetca_set_tuning(3, a.apply_stream)
    st = torch.cuda.current_stream().cuda_stream
    B = a.B
    shapes = SHAPES if not a.shapes else [tuple(int(v) for v in s.split(",")) for s in a.shapes]
    parts = torch.empty(lib.unetca_max_parts(B) * 4096, device="cuda")
    n = ctypes.c_int(0)
    bf = torch.bfloat16
    tot = {}
    for C, S in shapes:
        N = B * S * S * C
        y = torch.randn(B, S, S, C, device="cuda").to(bf)
        d = torch.randn(B, S, S, C, device="cuda").to(bf)
        out = torch.empty(B, S, S, 2 * C, device="cuda", dtype=bf)
        dy = torch.empty_like(y)
        pooled = torch.empty(B, S // 2, S // 2, C, device="cuda", dtype=bf)
        pos = torch.empty(B, S // 2, S // 2, C, device="cuda", dtype=torch.uint8)
        sc, sh = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
        mean, invstd = torch.randn(C, device="cuda") * 0.1, torch.rand(C, device="cuda") + 0.5
        s, dp = torch.rand(B, C, device="cuda"), torch.randn(B, C, device="cuda")
        coef = torch.rand(3, C, device="cuda")
        P = lambda t: t.data_ptr()  # noqa: E731
        kernels = {
            "bn_relu(write)     2N": (2 * N * 2, "ew", lambda: _lib.call("unetca_bn_relu", 1, P(y), C, P(dy), C, B, S * S, C, P(sc), P(sh), None, None, st)),
            "bn_relu(squeeze)   1N": (N * 2, "red", lambda: _lib.call("unetca_bn_relu", 1, P(y), C, None, 0, B, S * S, C, P(sc), P(sh), P(parts), ctypes.byref(n), st)),
            "se_squeeze(2 sums) 1N": (N * 2, "red", lambda: _lib.call("unetca_se_squeeze", 1, P(y), C, B, S * S, C, P(sc), P(sh), P(parts), ctypes.byref(n), st)),
            "se_scale_pool   2.25N": (2 * N * 2 + (N // 4) * 3, "fix", lambda: _lib.call("unetca_se_scale_pool", 1, P(y), C, P(out), 2 * C, P(pooled), C, P(pos), B, S, S, C, P(sc), P(sh), P(s), st)),
            "se_bn_bwd_reduce   2N": (2 * N * 2, "red", lambda: _lib.call("unetca_se_bn_bwd_reduce", 1, P(d), C, P(y), C, B, S * S, C, P(sc), P(sh), P(mean), P(parts), ctypes.byref(n), st)),
            "bn_bwd_reduce      2N": (2 * N * 2, "red", lambda: _lib.call("unetca_bn_bwd_reduce", 1, P(d), C, P(y), C, B, S * S, C, P(sc), P(sh), P(mean), P(invstd), None, None, P(parts), ctypes.byref(n), st)),
            "bn_bwd_apply(se)   3N": (3 * N * 2, "ew", lambda: _lib.call("unetca_bn_bwd_apply", 1, P(d), C, P(y), C, P(dy), C, B, S * S, C, P(sc), P(sh), P(mean), P(invstd), P(s), P(dp), P(coef), st)),
            "bn_bwd_apply       3N": (3 * N * 2, "ew", lambda: _lib.call("unetca_bn_bwd_apply", 1, P(d), C, P(y), C, P(dy), C, B, S * S, C, P(sc), P(sh), P(mean), P(invstd), None, None, P(coef), st)),
            "pool_bwd_add    2.25N": (2 * N * 2 + (N // 4) * 3, "fix", lambda: _lib.call("unetca_pool_bwd_add", 1, P(d), C, P(pooled), C, P(pos), P(dy), C, B, S, S, C, st)),
        }
        for name, (nbytes, kind, fn) in kernels.items():
            if a.only and a.only not in name:
                continue
            res = []
            settings = [(px, 1) for px in a.ew_px] if kind == "ew" else [(16, w) for w in a.red_waves] if kind == "red" else [(q, 1) for q in a.pool_quads]
            for px, w in settings:
                lib.unetca_set_tuning(0, px if kind != "fix" else 16)
                lib.unetca_set_tuning(1, w)
                if kind == "fix":
                    lib.unetca_set_tuning(2, px)
                ms = timeit(fn, a.iters)
                key = (name, w if kind == "red" else px)
                tot[key] = tot.get(key, 0.0) + ms
                res.append(f"{'waves' if kind == 'red' else 'px'}={w if kind == 'red' else px}: {ms:6.3f} ms {nbytes / ms / 1e6:6.0f} GB/s")
            print(f"C={C:4d} S={S:3d} {name}: " + " | ".join(res), flush=True)
        del y, d, out, dy, pooled, pos
    print("totals over the shapes (ms):")
    for (name, k), ms in tot.items():
        print(f"  {name} [{k}]: {ms:7.3f}")


if __name__ == "__main__":
    main()
