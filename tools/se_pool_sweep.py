"""BASELINE.json configs[4]: channel-attention + max-pool kernel sweep (C=64..1024, H=W=32..512) vs the HBM roofline,
with the plain max-pool (use_se=False ablation) beside it.  Calls the C ABI directly, times with CUDA events, tensor
>= 256 MB per case so the 126 MB L2 cannot hold it.  usage: python tools/se_pool_sweep.py [out.json]"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unetca_b200 import _lib  # noqa: E402


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def time_ms(fn, iters=10):
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def run_sweep(dtypes=("bf16", "f32"), iters=10, verbose=True):
    """-> (hbm peak GB/s, rows).  25 (C, HW) cells with B chosen so the tensor is >= 256 MB, plus the model's own five
    SE shapes at B = 64, for se_fwd / se_fwd+maxpool / maxpool_only (the use_se=False ablation)."""
    st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    peak = peak_gbs()
    rows = []
    for dt, tdt, e in ((_lib.BF16, torch.bfloat16, 2), (_lib.F32, torch.float32, 4)):
        if ("bf16" if e == 2 else "f32") not in dtypes:
            continue
        grid = [(C, S, max(1, -(-(256 << 20) // (C * S * S * e))), "cell") for C in (64, 128, 256, 512, 1024)
                for S in (32, 64, 128, 256, 512)]
        grid += [(C, S, 64, "model") for C, S in ((64, 512), (128, 256), (256, 128), (512, 64), (1024, 32))]   # the model's own layers
        for C, S, B, kind in grid:
            if B * C * S * S * e > (6 << 30):
                continue
            N = B * C * S * S
            y = torch.randn(B, S, S, C, device="cuda").to(tdt)
            out = torch.empty_like(y)
            pooled = torch.empty(B, S // 2, S // 2, C, dtype=tdt, device="cuda")
            pos = torch.empty(B, S // 2, S // 2, C, dtype=torch.uint8, device="cuda")
            scale, shift = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
            w1 = torch.randn(C // 16, C, device="cuda") / C ** 0.5
            w2 = torch.randn(C, C // 16, device="cuda") / (C // 16) ** 0.5
            parts = torch.empty(_lib.load().unetca_max_parts(B) * 4096, device="cuda")
            p_, z_, s_ = torch.empty(B, C, device="cuda"), torch.empty(B, C // 16, device="cuda"), torch.empty(B, C, device="cuda")
            n = ctypes.c_int(0)

            def se(pool):
                # the production sequence of model._double_conv_fwd: squeeze partial sums -> FC chain -> scale (+ pool)
                _lib.call("unetca_se_squeeze", dt, y.data_ptr(), C, B, S * S, C, scale.data_ptr(), shift.data_ptr(),
                          parts.data_ptr(), ctypes.byref(n), st())
                _lib.call("unetca_se_fc3", parts.data_ptr(), n.value, B, C, C // 16, S * S, w1.data_ptr(), w2.data_ptr(),
                          scale.data_ptr(), shift.data_ptr(), None, p_.data_ptr(), z_.data_ptr(), s_.data_ptr(), None, st())
                _lib.call("unetca_se_scale_pool", dt, y.data_ptr(), C, out.data_ptr(), C, pooled.data_ptr() if pool else None,
                          C if pool else 0, pos.data_ptr() if pool else None, B, S, S, C, scale.data_ptr(), shift.data_ptr(),
                          s_.data_ptr(), st())

            def pool_only():
                _lib.call("unetca_maxpool2x2", dt, y.data_ptr(), C, pooled.data_ptr(), C, pos.data_ptr(), None, B, S, S, C, st())

            cases = (("se_fwd", lambda: se(False), 3 * N * e),
                     ("se_fwd+maxpool", lambda: se(True), 3 * N * e + (N // 4) * (e + 1)),
                     ("maxpool_only", pool_only, N * e + (N // 4) * (e + 1)))
            for name, fn, nbytes in cases:
                ms = time_ms(fn, iters)
                gbs = nbytes / ms / 1e6
                rows.append({"dtype": "bf16" if e == 2 else "f32", "C": C, "HW": S, "B": B, "kind": kind, "kernel": name,
                             "ms": ms, "algorithmic_bytes": nbytes, "GBps": gbs, "frac_of_hbm_peak": gbs / peak})
                if verbose:
                    print(f"{rows[-1]['dtype']:4s} C={C:4d} HW={S:3d} B={B:4d} {name:16s} {ms:7.3f} ms {gbs:7.0f} GB/s "
                          f"{100 * gbs / peak:5.1f}% of {peak:.0f}", flush=True)
            del y, out, pooled, pos
    return peak, rows


def summarize(rows, dtype="bf16"):
    """min / median fraction of the HBM peak per kernel over the 25 cells, and the model's own five shapes at B = 64."""
    import statistics
    out = {}
    for k in ("se_fwd", "se_fwd+maxpool", "maxpool_only"):
        cells = [r["frac_of_hbm_peak"] for r in rows if r["dtype"] == dtype and r["kernel"] == k and r.get("kind") == "cell"]
        model = {f"C{r['C']}@{r['HW']}": round(r["frac_of_hbm_peak"], 3) for r in rows
                 if r["dtype"] == dtype and r["kernel"] == k and r.get("kind") == "model"}
        if cells:
            out[k] = {"cells": len(cells), "min_frac": round(min(cells), 3), "median_frac": round(statistics.median(cells), 3),
                      "model_shapes_B64": model}
    return out


def main():
    peak, rows = run_sweep()
    print(json.dumps(summarize(rows)))
    if len(sys.argv) > 1:
        json.dump({"hbm_peak_gbs": peak, "summary_bf16": summarize(rows), "rows": rows}, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
