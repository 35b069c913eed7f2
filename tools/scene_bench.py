"""BASELINE.json configs[3]: sliding-window inference on a synthetic 16384 x 16384 interferogram scene, spatial tiles
with halo, round-robin over the ranks of one box (no collective on the data path).

  python tools/scene_bench.py [--scene 16384] [--core 1024] [--halo 128] [--batch 2] [--precision bf16]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/scene_bench.py

The scene is never materialised: pixel (c, y, x) is a counter-based hash of its coordinates, so every rank generates
exactly the windows it needs (3.2 GB of fp32 scene would otherwise sit on every rank) and overlapping halos agree.
Prints one JSON line (rank 0): Mpx/s and tiles/s over all ranks, timed with CUDA events, max over ranks."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unetca_b200  # noqa: E402
from unetca_b200 import tiling  # noqa: E402


def window(tile, spec, H, W, dev):
    """(3, size, size) fp32 window of the hashed scene around `tile`, zero outside the scene."""
    ys = torch.arange(tile.y0 - spec.halo, tile.y0 - spec.halo + spec.size, device=dev, dtype=torch.int64)
    xs = torch.arange(tile.x0 - spec.halo, tile.x0 - spec.halo + spec.size, device=dev, dtype=torch.int64)
    c = torch.arange(3, device=dev, dtype=torch.int64)
    h = (ys[None, :, None] * 73856093) ^ (xs[None, None, :] * 19349663) ^ ((c[:, None, None] + 1) * 83492791)
    h = (h * 2654435761) & 0xFFFFFFFF
    h = ((h ^ (h >> 15)) * 2246822519) & 0xFFFFFFFF
    h = h ^ (h >> 13)
    v = (h & 0xFFFFFF).to(torch.float32) / float(1 << 23) - 1.0          # uniform [-1, 1)
    inside = ((ys >= 0) & (ys < H))[None, :, None] & ((xs >= 0) & (xs < W))[None, None, :]
    return torch.where(inside, v * 1.7320508, torch.zeros((), device=dev))


def run_scene(scene, core, halo, batch, precision, repeat, dev, rank=0, world=1):
    """Tiled eval-mode inference of this rank's share of a hashed scene x scene image.  -> (ms of the last pass, number
    of class-1 pixels, tiles of this rank, tiles in total, window size)."""
    torch.manual_seed(0)
    model = unetca_b200.UNet(3, 2, use_se=True).to(dev).set_precision(precision)
    spec = tiling.TileSpec(core=core, halo=halo)
    H = W = scene
    # random-init weights: give the BatchNorm running statistics real values (a few train-mode forwards on scene
    # windows, untimed) so that the eval-mode masks are not degenerate
    model.train()
    with torch.no_grad():
        for k in range(8):
            t0 = tiling.Tile(k, core * (k % max(1, scene // core)), core * (k % max(1, scene // core)), core, core)
            wnd = window(t0, spec, H, W, dev)
            model(torch.stack([wnd[:, :512, :512], wnd[:, 512:1024, 512:1024]]))
    model.eval()
    # ... and centre the two class logits (random weights leave a constant offset, i.e. an all-zero mask): shift the class-1
    # bias by the median logit difference of one window, the same on every rank (same seed, same window)
    with torch.no_grad():
        lg = model(window(tiling.Tile(0, 0, 0, core, core), spec, H, W, dev)[None, :, :512, :512])
        model.outc.bias.data[1] += (lg[:, 0] - lg[:, 1]).median()
    tiles = tiling.shard(tiling.plan(H, W, spec), rank, world)
    out = torch.empty(len(tiles), core, core, dtype=torch.uint8, device=dev)     # this rank's cores

    def one_pass():
        k = 0
        for t, m in tiling.predict_tiles(model, tiles, spec, lambda tl: window(tl, spec, H, W, dev), batch):
            out[k, :t.h, :t.w] = m
            k += 1

    ms = None
    for it in range(repeat):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_pass()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    ones = int((out == 1).sum().item())
    return ms, ones, len(tiles), len(tiling.plan(H, W, spec)), spec.size


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", type=int, default=16384)
    ap.add_argument("--core", type=int, default=1024)
    ap.add_argument("--halo", type=int, default=128)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--repeat", type=int, default=2, help="passes over this rank's tiles (the first is the warm-up)")
    ap.add_argument("--no-squeeze-fuse", action="store_true", help="SE squeeze as a separate pass (A/B of the epilogue fusion)")
    a = ap.parse_args()
    if a.no_squeeze_fuse:
        from unetca_b200 import model as _m
        _m.FUSE_SQUEEZE = False
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    H = W = a.scene
    ms, ones, my_tiles, ntiles, size = run_scene(a.scene, a.core, a.halo, a.batch, a.precision, a.repeat, dev, rank, world)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        t = torch.tensor([ones], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        ones = int(t.item())
    if rank == 0:
        print(json.dumps({
            "metric": "unet_ca_tiled_inference_megapixels_per_sec", "value": H * W / 1e6 / (ms / 1e3), "unit": "Mpx/s",
            "tiles_per_sec": ntiles / (ms / 1e3), "n_gpus": world, "ms": ms, "dtype": a.precision, "data": "synthetic",
            "config": {"workload": f"BASELINE configs[3]: {H}x{W} scene, core {a.core} + halo {a.halo} -> {size}^2 windows, "
                                   f"{ntiles} tiles round-robin over {world} GPU(s), {a.batch} tiles per forward, eval mode, "
                                   "window generation (coordinate hash) inside the timed region",
                       "tiles_per_rank": my_tiles},
            "class1_pixels": ones,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
