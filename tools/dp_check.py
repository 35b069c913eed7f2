"""Data-parallel parity check, run under torchrun with N >= 2 GPUs of one box:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
Every rank runs its shard of a global batch through the CUDA path with the bucketed, overlapped NCCL all-reduce
(unetca_b200.parallel.GradBuckets) and compares the resulting parameter gradients with the mean of the per-shard
gradients of the CPU oracle — fp32 mode 1e-2 on every gradient norm, bf16 1e-2 on the global norm —
and, to 1e-5, with the mean of the per-shard gradients of the same CUDA path run without data parallelism.
--microsteps k (BASELINE configs[2] mechanics): every rank runs k micro-batches with (loss / k).backward(), GradBuckets
accumulates locally and all-reduces on the k-th backward only; the oracle is the mean over all world*k micro-batches."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unetca_b200  # noqa: E402
from unetca_b200 import parallel  # noqa: E402
from oracle import unet_ca_port as port  # noqa: E402  (checker only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--microsteps", type=int, default=1)
    k = ap.parse_args().microsteps
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    per, H, W = 2, 64, 64            # per-rank shard; 64x64 keeps the bottleneck BatchNorm at 32 values per channel
    sd = port.make_state_dict(seed=9)
    nshard = world * k                                         # micro-batch (r, j) is shard r*k + j
    x, y = port.make_batch(9, per * nshard, H, W)
    ref = None
    for r in range(nshard):                                    # oracle: reference per shard, averaged on host
        _, _, g, _, _ = port.train_step_grads(sd, x[r * per:(r + 1) * per], y[r * per:(r + 1) * per])
        ref = g if ref is None else {n: ref[n] + g[n] for n in g}
    ref = {n: v / nshard for n, v in ref.items()}
    ok = True
    for prec in ("fp32", "bf16"):
        torch.manual_seed(1000 + rank)                         # deliberately different init: broadcast must fix it
        m = unetca_b200.UNet(3, 2, True).cuda().set_precision(prec)
        if rank == 0:
            m.load_state_dict(sd)
        gb = parallel.GradBuckets(m).set_accumulation(k)
        m.train()
        xs, ys = parallel.shard_batch(x, rank, world).cuda(), parallel.shard_batch(y, rank, world).cuda()
        for j in range(k):
            (m.loss(xs[j * per:(j + 1) * per], ys[j * per:(j + 1) * per]) / k).backward()
        torch.cuda.synchronize()
        params = dict(m.named_parameters())
        tot = torch.sqrt(sum((p.grad.float().cpu() ** 2).sum() for p in params.values())).item()
        rtot = torch.sqrt(sum((v ** 2).sum() for v in ref.values())).item()
        worst, worst_name = 0.0, ""
        for n, v in ref.items():
            if v.norm() > 1e-6 * rtot:
                e = abs(params[n].grad.float().cpu().norm().item() - v.norm().item()) / v.norm().item()
                if e > worst:
                    worst, worst_name = e, n
        # the data-parallel machinery itself, free of precision effects: the all-reduced gradient must equal the mean of
        # the per-shard gradients of the SAME CUDA path run shard by shard on this GPU without GradBuckets
        own = None
        for r in range(nshard):
            m2 = unetca_b200.UNet(3, 2, True).cuda().set_precision(prec)
            m2.load_state_dict(sd)
            m2.train()
            m2.loss(x[r * per:(r + 1) * per].cuda(), y[r * per:(r + 1) * per].cuda()).backward()
            g2 = {n: q.grad.double() for n, q in m2.named_parameters()}
            own = g2 if own is None else {n: own[n] + g2[n] for n in g2}
        dp_err = max(((params[n].grad.double() - own[n] / nshard).norm() / (own[n] / nshard).norm().clamp_min(1e-30)).item()
                     for n in own)
        # against the fp32 CPU oracle: every gradient norm in fp32 mode; in bf16 mode the global norm (the per-parameter
        # figure is reported: the tiny SE fc gradients of 2-image shards are dominated by bf16 rounding, identically so
        # with and without data parallelism)
        good = dp_err < 1e-5 and abs(tot - rtot) / rtot < 1e-2 and (worst < 1e-2 or prec != "fp32")
        # every rank must hold identical averaged gradients
        probe = torch.stack([p.grad.flatten()[0] for p in params.values()])
        gathered = [torch.zeros_like(probe) for _ in range(world)]
        dist.all_gather(gathered, probe)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        print(f"[rank {rank}] {prec} microsteps={k}: DP vs own mean-of-shards {dp_err:.2e}; global grad norm {tot:.6f} vs oracle mean-of-shards "
              f"{rtot:.6f}, worst per-param norm rel err {worst:.3e} ({worst_name}), identical across ranks: {same} -> "
              f"{'OK' if good and same else 'FAIL'}", flush=True)
        ok = ok and good and same
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
