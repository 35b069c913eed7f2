"""Micro-benchmark of the tcgen05 contraction kernels on one U-Net-CA layer shape (C ABI, CUDA events).
usage: python tools/conv_bench.py [--layers all|C,O,S ...] [--B 64] [--what fwd,wgrad] [--iters 10] [--block-n N]"""
import argparse
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unetca_b200 import _lib  # noqa: E402

LAYERS = [(64, 64, 512), (128, 64, 512), (64, 128, 256), (128, 128, 256), (256, 128, 256), (128, 256, 128),
          (256, 256, 128), (512, 256, 128), (256, 512, 64), (512, 512, 64), (1024, 512, 64), (512, 1024, 32),
          (1024, 1024, 32)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", nargs="*", default=["all"])
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--what", default="fwd,wgrad")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--block-n", type=int, default=0)
    ap.add_argument("--no-halo", action="store_true", help="narrow layers through the generic one-box-per-tap kernel")
    ap.add_argument("--kw64", action="store_true", help="64->64 layers through the kw-stacked kernel instead of the row-pair one")
    ap.add_argument("--convT", action="store_true", help="benchmark the four ConvTranspose2d(k2,s2) layers instead")
    ap.add_argument("--convT-narrow", action="store_true", help="ConvTranspose N blocks confined to one sub-pixel map (A/B)")
    ap.add_argument("--convT-generic", action="store_true", help="ConvTranspose forward through the generic pixels-on-M kernel (A/B)")
    ap.add_argument("--pixn-cluster", type=int, default=2, help="CTAs per cluster sharing weights by TMA multicast (1|2)")
    ap.add_argument("--wgrad-mode", type=int, default=0, help="0 auto, 1 narrow kernel everywhere, 2 no 256-wide tap-pair kernel")
    ap.add_argument("--no-kw", action="store_true", help="64-output layers through the row-pair kernel instead of the kw-stacked one")
    ap.add_argument("--no-pixn", action="store_true", help="narrow layers through the pixels-on-M generic kernel")
    ap.add_argument("--dgrad", action="store_true", help="add the dgrad shapes (O -> C) of the narrow layers")
    ap.add_argument("--rp64", action="store_true", help="64->64 layers through the resident-filter row-pair kernel")
    ap.add_argument("--bnstats", action="store_true", help="'fwd' = dgrad with the fused BatchNorm-backward statistics epilogue (C == O)")
    a = ap.parse_args()
    lib = _lib.load()
    lib.unetca_tc_force_block_n(a.block_n)
    lib.unetca_tc_force_no_pixn(1 if a.no_pixn else 0)
    lib.unetca_tc_force_wgrad_narrow(a.wgrad_mode)
    lib.unetca_tc_force_no_halo(1 if a.no_halo else 0)
    lib.unetca_tc_set_pixn_cluster(a.pixn_cluster)
    lib.unetca_tc_set_convT_wide(0 if a.convT_narrow else 1)
    lib.unetca_tc_set_convT_pix(0 if a.convT_generic else 1)
    lib.unetca_tc_set_convT_wgrad256(0 if a.convT_generic else 1)
    layers = LAYERS if a.layers == ["all"] else [tuple(int(v) for v in s.split(",")) for s in a.layers]
    if a.dgrad:
        layers = layers + [(64, 128, 512), (128, 64, 256), (256, 128, 128)]
    st = torch.cuda.current_stream().cuda_stream
    B = a.B
    ws = torch.empty(48 << 20, device="cuda")
    parts = torch.empty(lib.unetca_max_parts(B) * 2048, device="cuda")
    n = ctypes.c_int(0)
    if a.convT:
        for Cin, h in ((1024, 32), (512, 64), (256, 128), (128, 256)):
            Cout = Cin // 2
            x = torch.randn(B, h, h, Cin, device="cuda").bfloat16()
            cat = torch.empty(B, 2 * h, 2 * h, 2 * Cout, device="cuda", dtype=torch.bfloat16)
            dcat = torch.randn(B, 2 * h, 2 * h, 2 * Cout, device="cuda").bfloat16()
            wf = (torch.randn(4 * Cout, Cin, device="cuda") / Cin ** 0.5).bfloat16()
            wd = (torch.randn(Cin, 4 * Cout, device="cuda") / Cin ** 0.5).bfloat16()
            bias = torch.randn(Cout, device="cuda")
            dx = torch.empty_like(x)
            dw = torch.empty(Cin, Cout, 2, 2, device="cuda")
            fl = 2.0 * B * h * h * Cin * 4 * Cout
            up, dup = cat[..., Cout:], dcat[..., Cout:]
            fns = {
                "fwd": lambda: _lib.call("unetca_convT2x2_fwd", 1, x.data_ptr(), Cin, wf.data_ptr(), bias.data_ptr(), up.data_ptr(),
                                         2 * Cout, B, h, h, Cin, Cout, st),
                "dgrad": lambda: _lib.call("unetca_convT2x2_dgrad", 1, dup.data_ptr(), 2 * Cout, wd.data_ptr(), dx.data_ptr(), Cin, B,
                                           h, h, Cin, Cout, st),
                "wgrad": lambda: _lib.call("unetca_convT2x2_wgrad", 1, x.data_ptr(), Cin, dup.data_ptr(), 2 * Cout, ws.data_ptr(),
                                           ws.numel(), B, h, h, Cin, Cout, dw.data_ptr(), st),
            }
            res = []
            for what, fn in fns.items():
                for _ in range(3):
                    fn()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                s.record()
                for _ in range(a.iters):
                    fn()
                e.record()
                torch.cuda.synchronize()
                ms = s.elapsed_time(e) / a.iters
                res.append(f"{what} {ms:7.3f} ms {fl / ms / 1e9:7.1f} TFLOP/s")
            print(f"convT {Cin:4d}->{Cout:4d} @{h:3d}->{2 * h:3d} B={B}: " + " | ".join(res), flush=True)
            del x, cat, dcat
        return
    for C, O, S in layers:
        x = torch.randn(B, S, S, C, device="cuda").bfloat16()
        dy = torch.randn(B, S, S, O, device="cuda").bfloat16()
        y = torch.empty(B, S, S, O, device="cuda", dtype=torch.bfloat16)
        wf = (torch.randn(O, 9 * C, device="cuda") / (9 * C) ** 0.5).bfloat16()
        dw = torch.empty(O, C, 3, 3, device="cuda")
        fl = 2.0 * B * S * S * 9 * C * O
        wfp = wkw = None
        if O == 64 and (C == 128 or (C == 64 and a.kw64)) and not a.no_kw:
            wkw = torch.empty(9 * C, 64, device="cuda", dtype=torch.bfloat16)
            _lib.call("unetca_pack_conv3x3_kw", 1, wf.data_ptr(), 9 * C, wkw.data_ptr(), C, st)
        elif O % 128 and not a.no_pixn:
            wfp = torch.empty(2 * O, 12 * C, device="cuda", dtype=torch.bfloat16)
            _lib.call("unetca_pack_conv3x3_pair", 1, wf.data_ptr(), 9 * C, wfp.data_ptr(), O, C, st)
        res = []
        y1 = torch.randn(B, S, S, O, device="cuda").bfloat16() if a.bnstats else None
        cvec = [torch.rand(O, device="cuda") + 0.5, torch.randn(O, device="cuda") * 0.3, torch.randn(O, device="cuda") * 0.1]
        for what in a.what.split(","):
            if what == "fwd" and a.bnstats and C == O:
                fn = lambda: _lib.call("unetca_conv3x3_dgrad_bnstats", 1, x.data_ptr(), C, wf.data_ptr(), 9 * C, y.data_ptr(), O, B, S, S,
                                       C, O, y1.data_ptr(), O, cvec[0].data_ptr(), cvec[1].data_ptr(), cvec[2].data_ptr(),
                                       parts.data_ptr(), ctypes.byref(n), st)  # noqa: E731
            elif what == "fwd" and a.rp64 and C == 64 and O == 64:
                fn = lambda: _lib.call("unetca_conv3x3_fwd_rp64", 1, x.data_ptr(), C, wf.data_ptr(), 9 * C, y.data_ptr(), O, B, S, S,
                                       parts.data_ptr(), ctypes.byref(n), st)  # noqa: E731
            elif what == "fwd" and wkw is not None:
                fn = lambda: _lib.call("unetca_conv3x3_fwd_kw", 1, x.data_ptr(), C, wkw.data_ptr(), y.data_ptr(), O, B, S, S, C,
                                       parts.data_ptr(), ctypes.byref(n), st)  # noqa: E731
            elif what == "fwd" and wfp is not None:
                fn = lambda: _lib.call("unetca_conv3x3_fwd_paired", 1, x.data_ptr(), C, wfp.data_ptr(), y.data_ptr(), O, B, S, S,
                                       C, O, parts.data_ptr(), ctypes.byref(n), st)  # noqa: E731
            elif what == "fwd":
                fn = lambda: _lib.call("unetca_conv3x3_fwd", 1, x.data_ptr(), C, wf.data_ptr(), 9 * C, y.data_ptr(), O, B, S, S,
                                       C, O, parts.data_ptr(), ctypes.byref(n), st)  # noqa: E731
            else:
                fn = lambda: _lib.call("unetca_conv3x3_wgrad", 1, dy.data_ptr(), O, x.data_ptr(), C, ws.data_ptr(), ws.numel(),
                                       B, S, S, C, O, dw.data_ptr(), st)  # noqa: E731
            for _ in range(3):
                fn()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            s.record()
            for _ in range(a.iters):
                fn()
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / a.iters
            res.append(f"{what} {ms:7.3f} ms {fl / ms / 1e9:7.1f} TFLOP/s")
        print(f"conv {C:4d}->{O:4d} @{S:3d} B={B}: " + " | ".join(res), flush=True)
        del x, dy, y


if __name__ == "__main__":
    main()
