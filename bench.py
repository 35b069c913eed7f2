#!/usr/bin/env python
"""bench.py — U-Net-CA training throughput on B200 (BASELINE.json: "U-Net-CA train img/s @512^2 bf16").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one pass of the hot path over one batch of synthetic input: forward + softmax-CE loss + backward through
the drop-in `unetca_b200.UNet` (C ABI -> sm_100a kernels), the bucketed gradient all-reduce when N > 1, and
torch.optim.Adam(lr=1e-4) — the reference's train step (Unet-ChannalAttention.py:338-346).  Workload at every N
(weak scaling): BASELINE.json configs[1], U-Net-CA (use_se=True), bf16 mode, batch 64 per GPU, 3x512x512, 2 classes.

One JSON line on stdout (rank 0).  `value` = images/s with the batch resident in HBM, timed with NO per-kernel
instrumentation; `e2e` = the same step fed from pinned host memory (H2D copy of images+masks every step, prefetched
on a side stream, and the loss read back every step); `roofline` = the dominant tcgen05 contraction kernel's
algorithmic FLOP/s against the measured bf16 peak and `roofline_hbm` = the same for the memory-bound kernels against
the measured copy bandwidth, both from a SEPARATE pass of the same run in which every C-ABI call is bracketed by CUDA
events (the bracketing costs ~1 % and therefore stays out of `value`); `cpu_baseline` = the unmodified reference
classes timed on this box's host cores on a bounded sample.  `extra` carries the other BASELINE configs on the same
build: configs[2] (global batch 512 as 64-image micro-steps), the data-parallel gradient parity at N > 1, configs[3]
(tiled scene inference), configs[4] (SE + max-pool sweep, plain U-Net ablation) and the cuDNN arm.

--impl reference: the reference's own CPU implementation of the same step — `UNet` / `nn.CrossEntropyLoss` of the
UNMODIFIED Unet-ChannalAttention.py, which `__graft_entry__.build()` copies into the git-ignored oracle/_ref/ so that it
travels to the GPU box (kind "reference"; the oracle port stands in only if that copy is absent, kind "port") — on all
host threads, each step a bounded sample of the workload (2 images of 3x512x512).
"""
from __future__ import annotations

import argparse
import contextlib
import hashlib
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "unet_ca_train_images_per_sec_512x512_bf16"
UNIT = "img/s"
TRAIN_GFLOP_PER_IMG_512 = 1155.21       # BASELINE.md §4 (true shapes, no padding)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# per-kernel-class accounting: algorithmic FLOPs / bytes of every C-ABI call, CUDA events around the launches
# ------------------------------------------------------------------------------------------------------------
def _work(name, a, e, cin):
    """(class, algorithmic flops, algorithmic bytes) of one call; e = bytes per activation element."""
    if name == "unetca_conv3x3_fwd_kw":
        B, H, W, C = a[6:10]
        return "tensor", 2.0 * B * H * W * 9 * C * 64, 0
    if name == "unetca_conv3x3_fwd_paired":
        B, H, W, C, O = a[6:11]
        return "tensor", 2.0 * B * H * W * 9 * C * O, 0
    if name == "unetca_conv3x3_fwd_rp64":
        B, H, W = a[7:10]
        return "tensor", 2.0 * B * H * W * 9 * 64 * 64, 0
    if name == "unetca_conv3x3_fwd_split":
        B, H, W, C, O = a[10:15]
        return "tensor", 2.0 * B * H * W * 9 * C * O, 0
    if name == "unetca_conv3x3_dgrad_bnstats":
        B, H, W, C, O = a[7:12]
        return "tensor", 2.0 * B * H * W * 9 * C * O, 0
    if name == "unetca_conv3x3_fwd_cat":
        B, H, W, C, O = a[9:14]
        return "tensor", 2.0 * B * H * W * 9 * C * O, 0
    if name == "unetca_conv3x3_wgrad_cat":
        B, H, W, C, O = a[10:15]
        return "tensor", 2.0 * B * H * W * 9 * C * O, 0
    if name == "unetca_se_squeeze":
        B, hw, C = a[3:6]
        return "hbm", 0, B * hw * C * e
    if name == "unetca_conv3x3_fwd":
        B, H, W, C, O = a[7:12]
        return "tensor", 2.0 * B * H * W * 9 * C * O, 0
    if name == "unetca_conv3x3_wgrad":
        B, H, W, C, O = a[7:12]
        return "tensor", 2.0 * B * H * W * 9 * C * O, 0
    if name == "unetca_first_pairs_fwd":
        B, H, W, O = a[5:9]
        return "tensor", 2.0 * B * H * W * O * 9 * cin, 0
    if name == "unetca_first_pairs_wgrad":
        B, H, W, Cin, O = a[6:11]
        return "tensor", 2.0 * B * H * W * 9 * Cin * O, 0
    if name == "unetca_im2col_pairs":
        B, Cin, H, W = a[3:7]
        return "hbm", 0, B * H * W * Cin * 4 + B * (H // 2) * W * 64 * e
    if name == "unetca_gemm_nt":
        M, N = a[7], a[8]
        return "tensor", 2.0 * M * N * 9 * cin, 0
    if name == "unetca_im2col_wgrad":
        npix, Cin, O = a[7:10]
        return "tensor", 2.0 * npix * 9 * Cin * O, 0
    if name in ("unetca_convT2x2_fwd", "unetca_convT2x2_dgrad"):
        off = 7 if name.endswith("fwd") else 6
        B, h, w, Cin, Cout = a[off:off + 5]
        return "tensor", 2.0 * B * h * w * Cin * 4 * Cout, 0
    if name == "unetca_convT2x2_wgrad":
        B, h, w, Cin, Cout = a[7:12]
        return "tensor", 2.0 * B * h * w * Cin * 4 * Cout, 0
    if name == "unetca_bn_relu":
        B, hw, C = a[5:8]
        n = B * hw * C
        return "hbm", 0, n * e * (2 if a[3] else 1)
    if name == "unetca_se_scale_pool":
        B, H, W, C = a[8:12]
        n = B * H * W * C
        return "hbm", 0, 2 * n * e + ((n // 4) * (e + 1) if a[5] else 0)
    if name in ("unetca_se_bwd_reduce", "unetca_bn_bwd_reduce", "unetca_se_bn_bwd_reduce"):
        B, hw, C = a[5:8]
        return "hbm", 0, 2 * B * hw * C * e
    if name == "unetca_se_bn_bwd_reduce_pool":
        B, H, W, C = a[8:12]
        n = B * H * W * C
        return "hbm", 0, 2 * n * e + (n // 4) * (e + 1)
    if name == "unetca_bn_bwd_apply_pool":
        B, H, W, C = a[10:14]
        n = B * H * W * C
        return "hbm", 0, 3 * n * e + (n // 4) * (e + 1)
    if name == "unetca_bn_bwd_apply":
        B, hw, C = a[7:10]
        return "hbm", 0, 3 * B * hw * C * e
    if name == "unetca_pool_bwd_add":
        B, H, W, C = a[8:12]
        n = B * H * W * C
        return "hbm", 0, 2 * n * e + (n // 4) * (e + 1)
    if name == "unetca_chan_sum":
        return "hbm", 0, a[4] * a[3] * e
    if name == "unetca_outc_fwd":
        C, nc, B, HW = a[3], a[6], a[8], a[9]
        return "hbm", 0, B * HW * (C * e + nc * 4)
    if name == "unetca_outc_bwd":
        C, nc, B, HW = a[7], a[9], a[10], a[11]
        return "hbm", 0, B * HW * (2 * C * e + nc * 4)
    if name == "unetca_cross_entropy":
        nc, B, HW = a[2], a[3], a[4]
        return "hbm", 0, B * HW * (nc * 4 * 2 + 8)
    if name == "unetca_im2col3x3_nchw":
        B, Cin, H, W, Kpad = a[3:8]
        return "hbm", 0, B * H * W * (Cin * 4 + Kpad * e)
    return "other", 0, 0


class KernelAccount:
    def __init__(self, e, cin):
        self.e, self.cin, self.records, self.enabled = e, cin, [], False

    @contextlib.contextmanager
    def __call__(self, name, args):
        if not self.enabled:
            yield
            return
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        yield
        t.record()
        self.records.append((name, args, s, t))

    def summary(self, by_shape=False):
        """Per entry point; by_shape: per (entry point, integer arguments) — one row per layer shape."""
        out = {}
        for name, args, s, t in self.records:
            cls, fl, by = _work(name, args, self.e, self.cin)
            key = name
            if by_shape:
                key = name + "(" + ",".join(str(a) for a in args[1:] if isinstance(a, int) and 0 < a <= 4096) + ")"
            d = out.setdefault(key, {"class": cls, "ms": 0.0, "flops": 0.0, "bytes": 0.0, "calls": 0})
            d["ms"] += s.elapsed_time(t); d["flops"] += fl; d["bytes"] += by; d["calls"] += 1
        return out


# ------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the UNMODIFIED reference classes on host cores, bounded sample
# ------------------------------------------------------------------------------------------------------------
REF_COPY = os.path.join(ROOT, "oracle", "_ref", "Unet-ChannalAttention.py")     # placed by __graft_entry__.build()


def load_reference_module():
    """The reference's own module (a byte-identical copy under the git-ignored oracle/_ref/), or None."""
    if not os.path.exists(REF_COPY):
        return None
    spec = importlib.util.spec_from_file_location("unet_ca_reference", REF_COPY)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)          # no import-time side effects: everything runs under main() (UCA:489)
    return mod


def cpu_reference_step_time(steps, warmup, size, sample_b=2):
    """Time the reference train step (UCA:338-346: zero_grad, model(images), criterion, backward, Adam.step, loss.item())
    on the host.  -> (per-step seconds, images per step, kind)."""
    from oracle import unet_ca_port as port        # fixtures (+ fallback implementation); bench.py's baseline legs only
    torch.set_num_threads(os.cpu_count() or 1)
    sd = port.make_state_dict(seed=0)
    x, y = port.make_batch(0, sample_b, size, size)
    ref = load_reference_module()
    if ref is not None:
        kind = "reference"
        model = ref.UNet(in_channels=3, num_classes=2, use_se=True)
        model.load_state_dict(sd)
        model.train()
        crit = torch.nn.CrossEntropyLoss(ignore_index=255)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)

        def fwd():
            return crit(model(x), y)
    else:
        kind = "port"
        p = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
             for k, v in sd.items()}
        opt = torch.optim.Adam([v for v in p.values() if v.requires_grad], lr=1e-4)

        def fwd():
            return port.loss_fn(port.unet_forward(x, p, train=True), y)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = fwd()
        loss.backward()
        opt.step()
        _ = loss.item()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, sample_b, kind


KIND_NOTE = {"reference": "unmodified UNet + nn.CrossEntropyLoss of Unet-ChannalAttention.py (oracle/_ref copy made by build())",
             "port": "oracle/unet_ca_port.py (same ATen CPU kernels; oracle/_ref copy of the reference absent)"}


def run_reference_gpu(batch, size, steps, warmup):
    """EXTRA, not one of the contract's arms: the reference's own module stack on the B200 through cuDNN,
    channels_last + autocast(bf16) + fused Adam — the honest GPU competitor named in SURVEY.md 8(d)."""
    from oracle import unet_ca_port as port        # baseline leg only
    dev = torch.device("cuda", torch.cuda.current_device())
    B, S = batch, size
    sd = port.make_state_dict(seed=0)
    ref = load_reference_module()
    g = torch.Generator(device=dev).manual_seed(1234)
    x = torch.randn(B, 3, S, S, device=dev, generator=g).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 2, (B, S, S), device=dev, generator=g)
    torch.backends.cudnn.benchmark = True
    if ref is not None:
        kind = "reference"
        model = ref.UNet(in_channels=3, num_classes=2, use_se=True)
        model.load_state_dict(sd)
        model = model.to(dev).to(memory_format=torch.channels_last).train()
        crit = torch.nn.CrossEntropyLoss(ignore_index=255)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)

        def fwd():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                logits = model(x)
            return crit(logits.float(), y)
    else:
        kind = "port"
        p = {k: (v.to(dev).requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.to(dev))
             for k, v in sd.items()}
        opt = torch.optim.Adam([v for v in p.values() if v.requires_grad], lr=1e-4, fused=True)

        def fwd():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                logits = port.unet_forward(x, p, train=True)
            return port.loss_fn(logits.float(), y)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = fwd()
        loss.backward()
        opt.step()
        return loss

    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"impl": "reference-gpu (extra)", "metric": METRIC, "value": B * steps / (ms / 1e3), "unit": UNIT,
            "n_gpus": 1, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms / steps,
            "higher_is_better": True, "dtype": "bf16 autocast", "data": "synthetic", "kind": kind,
            "config": {"workload": f"the reference's module stack ({KIND_NOTE[kind]}) on one B200: cuDNN, channels_last, "
                                   f"autocast(bf16), fused Adam; batch {B}, 3x{S}x{S}"},
            "final_loss": loss.item()}


def config_dict(B, S, precision, world):
    """`config` of the JSON line — the same object on both arms (the reference arm runs a bounded sample of it)."""
    return {"workload": f"BASELINE configs[1]: U-Net-CA (use_se=True) train step fwd+CE+bwd+Adam(lr=1e-4), batch {B}/GPU, "
                        f"3x{S}x{S}, 2 classes, {precision} mode",
            "global_batch": B * world, "parallelism": f"dp{world}",
            "l2": "working set (~50 GB of activations per step) >> 126 MB L2; no explicit flush"}


def run_reference(args, rank):
    """Contract arm: the reference's CPU implementation on this box's host cores, on OUR arm's metric / unit / config;
    every step is a bounded sample (2 images) of that workload."""
    if rank != 0:
        return
    times, sb, kind = cpu_reference_step_time(args.steps, args.warmup, args.size)
    total = sum(times)
    v = sb * len(times) / total
    cores = torch.get_num_threads()
    sample = (f"{sb} images of 3x{args.size}x{args.size} per step (a bounded sample of the {args.batch}-image batch), fp32 "
              f"fwd+CE+bwd+Adam, {len(times)} timed steps after {args.warmup} warm-up; {KIND_NOTE[kind]}")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.batch, args.size, args.precision, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def csrc_hash():
    """Identity of the kernel sources a profile was captured on (tools/ncu_step_summary.py records it)."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "insar-unet-ca_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


# ------------------------------------------------------------------------------------------------------------
# extras: the other BASELINE configs on the same build (each guarded: a failure is recorded, never fatal)
# ------------------------------------------------------------------------------------------------------------
def guarded(fn):
    try:
        return fn()
    except Exception as e:                                       # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def record_timeline(step, x, y, rank, world, path):
    """One profiled step (all ranks run it, rank 0 records): every CUDA kernel with its stream, start and duration; the
    summary says how much NCCL time there is, how much of it runs under compute kernels, and how long the step's tail after
    the last compute kernel is."""
    import torch.distributed as dist
    from torch.profiler import ProfilerActivity, profile
    step(x, y)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if rank != 0:
        step(x, y)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return None
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(x, y)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ks = sorted(((e.name, e.time_range.start, e.time_range.end, getattr(e, "device_index", 0)) for e in ev), key=lambda t: t[1])
    if not ks:
        return {"error": "no CUDA kernel events recorded"}
    t0 = ks[0][1]
    nccl = [(a - t0, b - t0) for n, a, b, _ in ks if "nccl" in n.lower()]
    comp = [(n, a - t0, b - t0) for n, a, b, _ in ks if "nccl" not in n.lower() and "memcpy" not in n.lower() and "memset" not in n.lower()]

    def overlap(iv, others):
        tot = 0.0
        for a, b in iv:
            for c, d in others:
                lo, hi = max(a, c), min(b, d)
                if hi > lo:
                    tot += hi - lo
        return tot
    comp_iv = [(a, b) for _, a, b in comp]
    end_comp = max(b for _, _, b in comp)
    end_all = max(b for _, _, b, _ in ks) - t0
    # compute kernels that ran while an NCCL kernel was active, with their durations
    under = [(n, b - a) for n, a, b in comp if overlap([(a, b)], nccl) > 0.25 * (b - a)]
    # idle time of the compute stream: gaps between consecutive non-NCCL kernels (launch latency, host not far enough ahead)
    gaps = []
    cur_end = comp[0][2]
    for (n, a, b), (pn, _, _) in zip(comp[1:], comp[:-1]):
        if a > cur_end:
            gaps.append((a - cur_end, pn[:50], n[:50]))
        cur_end = max(cur_end, b)
    gaps.sort(reverse=True)
    out = {"idle_us_between_compute_kernels": sum(g[0] for g in gaps), "gaps_over_20us": len([g for g in gaps if g[0] > 20]),
           "largest_gaps": [{"us": round(g[0], 1), "after": g[1], "before": g[2]} for g in gaps[:25]],
           "step_us": end_all, "kernels": len(ks), "nccl_kernels": len(nccl), "nccl_us": sum(b - a for a, b in nccl),
           "nccl_us_under_compute": overlap(nccl, comp_iv), "tail_after_last_compute_us": end_all - end_comp,
           "compute_us": sum(b - a for a, b in comp_iv),
           "compute_kernels_running_under_nccl": [{"name": n[:60], "us": d} for n, d in under][:40],
           "nccl_intervals_us": [[round(a, 1), round(b, 1)] for a, b in nccl][:40]}
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    return {k: out[k] for k in ("step_us", "kernels", "idle_us_between_compute_kernels", "gaps_over_20us", "nccl_kernels", "nccl_us",
                                "nccl_us_under_compute", "tail_after_last_compute_us")}


def extra_dp_parity(dev, rank, world):
    """N > 1: one untimed tiny fp32 train step.  The all-reduced gradients that GradBuckets leaves in .grad against the
    all-gathered mean of the per-rank gradients of the SAME model run without data parallelism, and whether every
    rank ends up with bit-identical gradients."""
    import torch.distributed as dist
    import unetca_b200
    from unetca_b200 import parallel
    torch.manual_seed(4321)                                      # same init everywhere
    m = unetca_b200.UNet(3, 2, use_se=True).to(dev).set_precision("fp32")
    m.train()
    g = torch.Generator(device=dev).manual_seed(99 + rank)       # a different shard per rank
    x = torch.randn(2, 3, 64, 64, device=dev, generator=g)
    y = torch.randint(0, 2, (2, 64, 64), device=dev, generator=g)
    m.loss(x, y).backward()
    local = torch.cat([p.grad.flatten() for p in m.parameters()])
    for p in m.parameters():
        p.grad = None
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    want = torch.stack(gathered).double().mean(0)
    gb = parallel.GradBuckets(m)
    # BatchNorm running statistics moved in the first forward; the forward itself (batch statistics) does not read them
    m.loss(x, y).backward()
    got = torch.cat([p.grad.flatten() for p in m.parameters()]).double()
    gb.detach()
    err = ((got - want).norm() / want.norm()).item()
    worst = ((got - want).abs().max() / want.abs().max()).item()
    mine = got.float()
    all_g = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(all_g, mine)
    same = all(torch.equal(all_g[0], t) for t in all_g)
    return {"rel_l2_err_vs_mean_of_rank_grads": err, "max_abs_err_over_max": worst, "identical_across_ranks": bool(same),
            "params": int(got.numel()), "what": "fp32 mode, 2 images of 3x64x64 per rank, bucketed overlapped all-reduce vs "
                                               "all-gathered mean of the same model's per-rank gradients"}


def extra_global_batch(model, opt, buckets, x, y, world, rank, dev, global_batch=512, steps=3):
    """BASELINE configs[2]: global batch 512 at every N as k = 512 / (64 N) micro-steps of 64 images per GPU (the
    BatchNorm batch stays 64), gradients accumulated locally and all-reduced on the k-th backward only."""
    import torch.distributed as dist
    B = x.shape[0]
    k = max(1, global_batch // (B * world))
    if buckets is not None:
        buckets.set_accumulation(k)

    def step():
        opt.zero_grad(set_to_none=True)
        for _ in range(k):
            (model.loss(x, y) / k).backward()
        opt.step()

    step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    if buckets is not None:
        buckets.set_accumulation(1)
    return {"value": k * B * world * steps / (ms / 1e3), "unit": UNIT, "global_batch": k * B * world, "micro_steps": k,
            "images_per_gpu_per_micro_step": B, "optimizer_steps_timed": steps, "ms_per_optimizer_step": ms / steps,
            "scaling": "strong", "all_reduces_per_optimizer_step": 1 if world > 1 else 0}


def extra_rank_spread(model, opt, buckets, x, y, world, rank, dev, steps=10):
    """N > 1 diagnostic: every rank runs `steps` train steps AT THE SAME TIME with the all-reduce switched off (no_sync), so
    no rank waits for another, and reports its own time.  The spread is what the slowest-GPU-sets-the-pace part of the
    1 -> N loss looks like; the difference between the slowest rank here and the synchronised step is what the exchange
    itself costs."""
    import torch.distributed as dist

    def step():
        opt.zero_grad(set_to_none=True)
        model.loss(x, y).backward()
        opt.step()

    with buckets.no_sync():
        for _ in range(3):
            step()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    ms = sorted(v.item() for v in allt)
    return {"ms_per_step_by_rank_sorted": [round(v, 2) for v in ms], "min": ms[0], "max": ms[-1], "steps": steps,
            "what": "all ranks stepping simultaneously without the gradient all-reduce (no_sync): per-rank ms/step"}


def extra_ablation(dev, B, S, precision, steps=3):
    """configs[4]'s ablation baseline: the plain U-Net (use_se=False, /root/reference/Unet.py) through the same path."""
    import unetca_b200
    from unetca_b200 import optim as uoptim
    torch.manual_seed(0)
    m = unetca_b200.UNet(3, 2, use_se=False).to(dev).set_precision(precision)
    m.train()
    opt = uoptim.Adam(m.parameters(), lr=1e-4, model=m)
    g = torch.Generator(device=dev).manual_seed(77)
    x = torch.randn(B, 3, S, S, device=dev, generator=g)
    y = torch.randint(0, 2, (B, S, S), device=dev, generator=g)

    def step():
        opt.zero_grad(set_to_none=True)
        m.loss(x, y).backward()
        opt.step()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": B / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "workload": f"plain U-Net (use_se=False), batch {B}, 3x{S}x{S}"}


def extra_scene(dev, rank=0, world=1):
    """BASELINE configs[3]: tiled eval-mode inference, tiles round-robin over the ranks, no data-path collective.  One GPU: a
    4096^2 scene (16 tiles); 8 GPUs: the 16384^2 scene configs[3] names (256 tiles, 32 per rank); 2 / 4 GPUs: 8192^2."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import scene_bench
    scene = {1: 4096, 2: 8192, 4: 8192, 8: 16384}.get(world, 4096)
    ms, ones, my_tiles, ntiles, size = scene_bench.run_scene(scene, 1024, 128, 4, "bf16", 2, dev, rank, world)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        t = torch.tensor([ones], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        ones = int(t.item())
    return {"value": scene * scene / 1e6 / (ms / 1e3), "unit": "Mpx/s", "ms": ms, "tiles": ntiles, "tiles_per_rank": my_tiles,
            "tiles_per_sec": ntiles / (ms / 1e3), "n_gpus": world,
            "workload": f"BASELINE configs[3]{' at reduced extent' if scene != 16384 else ''}: {scene}x{scene} hashed scene, core 1024 + "
                        f"halo 128 -> {size}^2 windows, {ntiles} tiles round-robin over {world} GPU(s), 4 per forward, eval mode, bf16, "
                        "second pass timed (max over ranks)", "class1_pixels": ones}


def extra_se_sweep():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import se_pool_sweep
    peak, rows = se_pool_sweep.run_sweep(dtypes=("bf16",), iters=5, verbose=False)
    out = se_pool_sweep.summarize(rows)
    out["hbm_peak_gbs"] = peak
    out["workload"] = ("BASELINE configs[4], bf16: C in {64..1024} x HW in {32..512}, B sized for >= 256 MB tensors; fraction of the "
                       "measured HBM copy bandwidth at 3*N*e (+ pool) algorithmic bytes; full table: tools/se_pool_sweep.py")
    return out


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"],
                    help="reference = the reference's CPU path (contract arm); reference-gpu = EXTRA: the same modules on the "
                         "GPU through cuDNN (channels_last + autocast bf16)")
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the `extra` object (configs[2..4], DP parity, cuDNN arm)")
    ap.add_argument("--no-kernel-pass", action="store_true", help="skip the instrumented pass (no roofline objects)")
    ap.add_argument("--kernel-steps", type=int, default=5, help="steps of the instrumented pass (CUDA events around every call)")
    ap.add_argument("--torch-adam", action="store_true",
                    help="step with torch.optim.Adam(fused=True) instead of unetca_b200.optim.Adam (A/B)")
    ap.add_argument("--graph", action="store_true",
                    help="replay the whole train step as one CUDA graph (unetca_b200.graph; single GPU; no per-kernel "
                         "accounting, so the roofline objects are omitted) — what matters at small batch / tile sizes")
    ap.add_argument("--kernel-table", default=None, help="write the per-kernel table (JSON) to this path")
    ap.add_argument("--nccl-max-ctas", type=int, default=int(os.environ.get("UNETCA_NCCL_MAX_CTAS", "0")),
                    help="N > 1: cap the CTAs NCCL may use for the gradient all-reduce (0 = NCCL's default)")
    ap.add_argument("--compute-priority", type=int, default=int(os.environ.get("UNETCA_COMPUTE_PRIORITY", "0")),
                    help="A/B: run the step on a CUDA stream of this priority (-1 = above NCCL's stream) instead of the default stream")
    ap.add_argument("--timeline", default=None,
                    help="N > 1: rank 0 records one extra step with torch.profiler and writes a per-stream kernel timeline "
                         "summary (JSON) here: NCCL time, overlap with compute, exposed tail")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.impl == "reference-gpu":
        if rank == 0:
            torch.cuda.set_device(0)
            print(json.dumps(run_reference_gpu(args.batch, args.size, args.steps, args.warmup)), flush=True)
        return

    import torch.distributed as dist
    import unetca_b200
    from unetca_b200 import _lib, parallel
    from unetca_b200 import optim as uoptim

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        pg_opts = None
        if args.nccl_max_ctas > 0:
            # the bucket all-reduces run UNDER the persistent one-CTA-per-SM contraction kernels of backward: every SM NCCL
            # takes delays a CTA that owns a fixed share of tiles, so the collective is capped to a few CTAs (125 MB per
            # ~90 ms step needs < 5 GB/s of the 900 GB/s NVLink port)
            pg_opts = dist.ProcessGroupNCCL.Options()
            pg_opts.config.max_ctas = args.nccl_max_ctas
            pg_opts.config.min_ctas = min(args.nccl_max_ctas, 1)
        dist.init_process_group("nccl", device_id=dev, pg_options=pg_opts)
    if args.warmup < 3:
        args.warmup = 3
    if args.compute_priority:
        # the persistent contraction kernels want all 148 SMs at every kernel boundary; on a higher-priority stream their CTAs
        # are placed before the pending CTAs of a bucket all-reduce (which then runs in the shadow of the lighter kernels)
        torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=args.compute_priority))

    if os.environ.get("UNETCA_BLOCK_N"):                 # tuning knob: force the tcgen05 tile width where it divides N
        _lib.load().unetca_tc_force_block_n(int(os.environ["UNETCA_BLOCK_N"]))
    if os.environ.get("UNETCA_EW_STREAM"):               # tuning knob: 0 = register kernels for the BN-backward / squeeze passes
        _lib.load().unetca_set_tuning(3, int(os.environ["UNETCA_EW_STREAM"]))
    B, S = args.batch, args.size
    torch.manual_seed(0)
    model = unetca_b200.UNet(3, 2, use_se=True).to(dev).set_precision(args.precision)
    model.train()
    buckets = parallel.GradBuckets(model) if world > 1 else None
    if args.torch_adam:
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    else:
        # this repo's multi-tensor Adam: same arithmetic, emits the packed conv filters for the next forward
        opt = uoptim.Adam(model.parameters(), lr=1e-4, model=model)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn(B, 3, S, S, device=dev, generator=g)
    y = torch.randint(0, 2, (B, S, S), device=dev, generator=g)

    def step(xb, yb):
        opt.zero_grad(set_to_none=True)
        loss = model.loss(xb, yb)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.graph:
        if world > 1:
            raise SystemExit("bench.py --graph: single GPU only")
        from unetca_b200 import graph as ugraph
        if args.torch_adam:
            opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True)
        gstep = ugraph.GraphedTrainStep(model, opt, x, y, warmup=args.warmup)
        torch.cuda.synchronize()
        clocks = ClockSampler(local)
        clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = gstep(x, y)
        e1.record()
        torch.cuda.synchronize()
        clk = clocks.stop()
        ms = e0.elapsed_time(e1)
        print(json.dumps({"metric": METRIC, "value": B * args.steps / (ms / 1e3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                          "config": {"workload": f"U-Net-CA (use_se=True) train step fwd+CE+bwd+Adam(capturable) replayed as ONE CUDA "
                                                 f"graph, batch {B}, 3x{S}x{S}, {args.precision} mode"},
                          "clocks": clk, "final_loss": loss.item()}), flush=True)
        return

    for _ in range(args.warmup):
        step(x, y)
    # ---- timed region 1: batch resident in HBM; NO per-kernel instrumentation ---------------------------------
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    n0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step(x, y)
    e1.record()
    barrier()
    launches = _lib.launch_count - n0
    clk = clocks.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = world * B * args.steps / (ms / 1e3)
    ms_step = ms / args.steps
    final_loss = loss.item()

    # ---- timed region 2: end to end from pinned host memory ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        hx = [torch.randn(B, 3, S, S).pin_memory() for _ in range(2)]
        hy = [torch.randint(0, 2, (B, S, S)).pin_memory() for _ in range(2)]
        dx = [torch.empty_like(x) for _ in range(2)]
        dy = [torch.empty_like(y) for _ in range(2)]
        copy_stream = torch.cuda.Stream()
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]

        def upload(i):
            b = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[b])
                dx[b].copy_(hx[b], non_blocking=True)
                dy[b].copy_(hy[b], non_blocking=True)
                ready[b].record(copy_stream)

        # the step's result (the loss) is read back EVERY step: an async D2H copy into pinned memory right behind the step,
        # consumed by the host one step later (deferred logging) — the host keeps queueing the next step meanwhile instead of
        # draining the GPU at every `.item()`; all reads are complete inside the timed region
        hloss = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
        got = [torch.cuda.Event() for _ in range(2)]
        losses = []
        for b in range(2):
            free[b].record()
        barrier()
        t0 = time.perf_counter()
        upload(0)
        for i in range(args.steps):
            if i + 1 < args.steps:
                upload(i + 1)
            b = i % 2
            torch.cuda.current_stream().wait_event(ready[b])
            l = step(dx[b], dy[b])
            free[b].record()
            hloss[b].copy_(l.detach(), non_blocking=True)      # device -> host read of the step's result, every step
            got[b].record()
            if i >= 1:
                got[1 - b].synchronize()
                losses.append(float(hloss[1 - b]))
        got[(args.steps - 1) % 2].synchronize()
        losses.append(float(hloss[(args.steps - 1) % 2]))
        barrier()
        dt_e2e = time.perf_counter() - t0
        assert len(losses) == args.steps and all(v == v for v in losses)
        if world > 1:
            t = torch.tensor([dt_e2e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_e2e = t.item()
        e2e = {"value": world * B * args.steps / dt_e2e, "unit": UNIT,
               "h2d_bytes_per_step": int(hx[0].numel() * 4 + hy[0].numel() * 8), "d2h_bytes_per_step": 4}
        del hx, hy, dx, dy

    # ---- instrumented pass (same run, same tensors): CUDA events around every C-ABI call -----------------------
    roof = roof_all = roof_h = None
    pk = peaks()
    if not args.no_kernel_pass:
        acct = KernelAccount(2 if args.precision == "bf16" else 4, 3)
        _lib.set_hook(acct)
        acct.enabled = True
        barrier()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(args.kernel_steps):
            step(x, y)
        k1.record()
        barrier()
        acct.enabled = False
        _lib.set_hook(None)
        ms_kstep = k0.elapsed_time(k1) / args.kernel_steps
        table = acct.summary()
        nk = args.kernel_steps
        # DRAM traffic per launch (dram__bytes_read + dram__bytes_write) from the committed ncu capture of this command —
        # only if it was taken on the kernel sources that are running now
        traffic, traffic_note = {}, None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp) and B == 64 and S == 512 and args.precision == "bf16":
            tj = json.load(open(tp))
            if tj.get("csrc_sha") == csrc_hash():
                traffic = tj
                traffic_note = tj.get("source")
            else:
                traffic_note = (f"profiles/ncu_traffic.json was captured on kernel sources {tj.get('csrc_sha')}, this build is "
                                f"{csrc_hash()}: traffic withheld (re-run the ncu pass of tools/ncu_step_summary.py)")
        tms = sum(d["ms"] for d in table.values() if d["class"] == "tensor")
        tfl = sum(d["flops"] for d in table.values() if d["class"] == "tensor")
        hms = sum(d["ms"] for d in table.values() if d["class"] == "hbm")
        hby = sum(d["bytes"] for d in table.values() if d["class"] == "hbm")
        ncalls_t = sum(d["calls"] for d in table.values() if d["class"] == "tensor")
        ncalls_h = sum(d["calls"] for d in table.values() if d["class"] == "hbm")
        ach_t = tfl / (tms * 1e-3) / 1e12 if tms else 0.0
        ach_h = hby / (hms * 1e-3) / 1e9 if hms else 0.0

        def one(name):
            d = table.get(name)
            return d if d and d["ms"] else None

        def kernel_rows(pred):
            """Aggregate the records whose (entry point, args) satisfy pred — a KERNEL may sit behind several entry points."""
            d = {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "calls": 0}
            for name, a, s0, t0 in acct.records:
                if pred(name, a):
                    _, fl, by = _work(name, a, acct.e, acct.cin)
                    d["ms"] += s0.elapsed_time(t0); d["flops"] += fl; d["bytes"] += by; d["calls"] += 1
            return d if d["calls"] and d["ms"] else None

        how = (f"separate instrumented pass of the same run: {nk} steps with CUDA events around every C-ABI call "
               f"({ms_kstep:.2f} ms/step instrumented vs {ms_step:.2f} in the headline region)")
        # dominant kernel of the step: the haloed pixels-on-N conv kernel behind unetca_conv3x3_fwd (forward + dgrad of every
        # layer with O % 128 == 0); the aggregate over all contraction kernels is kept beside it
        # (tc_conv3x3_hpix_kernel: every unetca_conv3x3_fwd call here has O % 128 == 0, plus the dgrads with the fused
        # BatchNorm-backward statistics epilogue, unetca_conv3x3_dgrad_bnstats with O % 128 == 0)
        dom = kernel_rows(lambda n, a: n in ("unetca_conv3x3_fwd", "unetca_conv3x3_fwd_split") or
                          (n == "unetca_conv3x3_dgrad_bnstats" and a[11] % 128 == 0) or
                          (n == "unetca_conv3x3_fwd_cat" and a[13] % 128 == 0))
        roof_all = {"bound": "tensor", "achieved": ach_t, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach_t / pk["tflops_sustained"], "traffic": traffic.get("tensor", {}).get("dram_bytes_per_launch"),
                    "kernel": "all tcgen05 contraction kernels (conv3x3 fwd/dgrad/wgrad, ConvTranspose, first conv), aggregate",
                    "launches": ncalls_t, "avg_launch_ms": tms / max(ncalls_t, 1), "share_of_step": tms / nk / ms_kstep,
                    "flops_per_launch_avg": tfl / max(ncalls_t, 1), "peak_source": pk["source"] + " bf16_tflops_sustained"}
        if dom:
            ach_d = dom["flops"] / (dom["ms"] * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach_d, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach_d / pk["tflops_sustained"], "frac_of_burst_peak": ach_d / pk["tflops_burst"],
                    "traffic": traffic.get("dominant", {}).get("dram_bytes_per_launch"), "traffic_source": traffic_note,
                    "kernel": "tc_conv3x3_hpix_kernel (tcgen05 haloed pixels-on-N conv3x3 forward/dgrad; entries unetca_conv3x3_fwd, "
                              "unetca_conv3x3_fwd_split (two-destination epilogue), unetca_conv3x3_fwd_cat (two-source operand) and, with the fused BN-backward statistics epilogue, "
                              "unetca_conv3x3_dgrad_bnstats)",
                    "launches": dom["calls"], "avg_launch_ms": dom["ms"] / dom["calls"], "share_of_step": dom["ms"] / nk / ms_kstep,
                    "flops_per_launch_avg": dom["flops"] / dom["calls"], "peak_source": pk["source"] + " bf16_tflops_sustained",
                    "measured": how,
                    "algorithmic_note": "2*B*H*W*9*C*O per launch (true shapes); `peak` is the measured SUSTAINED cuBLAS bf16 rate "
                                        "of MEASURED_PEAKS.json (a frac near or above 1 means this kernel holds what cuBLAS holds "
                                        "under the same power cap; frac_of_burst_peak uses the burst figure)"}
        else:
            roof = roof_all
        domh = one("unetca_bn_bwd_apply")
        roof_h = {"bound": "hbm", "achieved": ach_h, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach_h / pk["hbm_gbs"],
                  "traffic": traffic.get("hbm", {}).get("dram_bytes_per_launch"),
                  "algorithmic_bytes_per_launch": hby / max(ncalls_h, 1),
                  "kernel": "BN/ReLU/SE/max-pool/outc/CE elementwise and reduction kernels (aggregate)",
                  "launches": ncalls_h, "share_of_step": hms / nk / ms_kstep, "peak_source": pk["source"] + " hbm_gbs"}
        if domh:
            ach = domh["bytes"] / (domh["ms"] * 1e-3) / 1e9
            roof_h["dominant"] = {"kernel": "bn_bwd_apply_stream_kernel (entry unetca_bn_bwd_apply: ReLU+BN backward as a cp.async.bulk shared-memory stream, 3*N*e bytes)",
                                  "achieved": ach, "frac": ach / pk["hbm_gbs"], "launches": domh["calls"],
                                  "avg_launch_ms": domh["ms"] / domh["calls"], "share_of_step": domh["ms"] / nk / ms_kstep,
                                  "algorithmic_bytes_per_launch": domh["bytes"] / domh["calls"],
                                  "traffic": traffic.get("dominant_hbm", {}).get("dram_bytes_per_launch")}
        if args.kernel_table and rank == 0:
            with open(args.kernel_table, "w") as f:
                json.dump({k: {**v, "ms_per_step": v["ms"] / nk} for k, v in table.items()}, f, indent=1)
            with open(args.kernel_table.replace(".json", "") + "_by_shape.json", "w") as f:
                json.dump({k: {**v, "ms_per_step": v["ms"] / nk} for k, v in acct.summary(by_shape=True).items()}, f, indent=1)

    if args.timeline:
        tl = guarded(lambda: record_timeline(step, x, y, rank, world, args.timeline))
        if rank == 0:
            print("timeline:", json.dumps(tl)[:600], file=sys.stderr)

    # ---- extras: the other BASELINE configs on this build ------------------------------------------------------
    extra = None
    if not args.no_extras:
        extra = {}
        extra["global_batch_512"] = guarded(lambda: extra_global_batch(model, opt, buckets, x, y, world, rank, dev))
        if world > 1:
            if buckets is not None:
                extra["rank_spread_no_sync"] = guarded(lambda: extra_rank_spread(model, opt, buckets, x, y, world, rank, dev))
                buckets.detach()
            extra["dp_parity"] = guarded(lambda: extra_dp_parity(dev, rank, world))
    del model, opt, buckets
    torch.cuda.empty_cache()
    if extra is not None and world > 1:
        extra["scene_inference"] = guarded(lambda: extra_scene(dev, rank, world))
        torch.cuda.empty_cache()
    if extra is not None and world == 1:
        extra["ablation_plain_unet"] = guarded(lambda: extra_ablation(dev, B, S, args.precision))
        torch.cuda.empty_cache()
        extra["scene_inference"] = guarded(lambda: extra_scene(dev))
        torch.cuda.empty_cache()
        extra["se_pool_sweep"] = guarded(extra_se_sweep)
        torch.cuda.empty_cache()
        del x, y
        extra["cudnn_arm"] = guarded(lambda: run_reference_gpu(B, S, 5, 3))
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        times, sb, kind = cpu_reference_step_time(5, 2, S)
        v = sb * len(times) / sum(times)
        # BASELINE.json configs[0] exactly: batch 4, 3x256x256, fp32, best of 3 after 1 warm-up
        t0, _, _ = cpu_reference_step_time(3, 1, 256, sample_b=4)
        cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
               "sample": f"{sb} images of 3x{S}x{S} per step (of the {B}-image batch), fp32 fwd+CE+bwd+Adam, {len(times)} timed "
                         f"steps after 2 warm-up; {KIND_NOTE[kind]}",
               "configs0": {"value": 4 / min(t0), "unit": UNIT,
                            "sample": "BASELINE configs[0]: batch 4, 3x256x256, fp32 fwd+CE+bwd+Adam, best of 3 after 1 warm-up"}}

    if rank == 0:
        gflop_img = TRAIN_GFLOP_PER_IMG_512 * (S * S) / (512 * 512)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": config_dict(B, S, args.precision, world),
            "tensor_util_step": value / world * gflop_img * 1e9 / (pk["tflops_sustained"] * 1e12),
            "roofline": roof, "roofline_tensor_all": roof_all, "roofline_hbm": roof_h, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clk, "final_loss": final_loss, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    # stdout carries exactly ONE JSON line: native libraries write to fd 1 as well (NCCL prints its version banner
    # there), so fd 1 is pointed at stderr and Python's own stdout keeps the original descriptor
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_real_stdout, "w")
    main()
    sys.stdout.flush()
