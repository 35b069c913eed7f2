"""Host-side logic of tiled inference: the plan partitions the scene, shards are a round-robin partition of the
plan, windows are zero-filled outside the scene."""
import pytest
import torch


def test_plan_partitions_scene_and_shards():
    from unetca_b200 import tiling
    spec = tiling.TileSpec(core=32, halo=16)
    H, W = 70, 100
    tiles = tiling.plan(H, W, spec)
    cover = torch.zeros(H, W, dtype=torch.int32)
    for t in tiles:
        cover[t.y0:t.y0 + t.h, t.x0:t.x0 + t.w] += 1
    assert torch.all(cover == 1)
    assert len(tiles) == 3 * 4
    shards = [tiling.shard(tiles, r, 8) for r in range(8)]
    assert sorted(t.index for s in shards for t in s) == list(range(len(tiles)))
    assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    # BASELINE configs[3]: 16384^2 scene, core 1024 + halo 128 -> 256 tiles of 1280^2, 32 per GPU
    big = tiling.plan(16384, 16384, tiling.TileSpec(1024, 128))
    assert len(big) == 256 and all(len(tiling.shard(big, r, 8)) == 32 for r in range(8))
    assert tiling.TileSpec(1024, 128).size == 1280


def test_extract_zero_fills_outside():
    from unetca_b200 import tiling
    spec = tiling.TileSpec(core=32, halo=16)
    scene = torch.arange(3 * 40 * 50, dtype=torch.float32).view(3, 40, 50) + 1
    t0, tl = tiling.plan(40, 50, spec)[0], tiling.plan(40, 50, spec)[-1]
    w = tiling.extract(scene, t0, spec)
    assert w.shape == (3, 64, 64)
    assert torch.all(w[:, :16, :] == 0) and torch.all(w[:, :, :16] == 0)
    assert torch.equal(w[:, 16:56, 16:64], scene[:, :40, :48])
    w = tiling.extract(scene, tl, spec)                      # core at (32, 32), extent 8 x 18
    assert torch.equal(w[:, 16:24, 16:34], scene[:, 32:40, 32:50])
    assert torch.all(w[:, 24:, :] == 0) and torch.all(w[:, :, 34:] == 0)


def test_tilespec_validation():
    from unetca_b200 import tiling
    with pytest.raises(ValueError):
        tiling.TileSpec(core=8, halo=2)                       # 12 < 16: four 2x2 poolings impossible
    with pytest.raises(ValueError):
        tiling.TileSpec(core=0, halo=8)
    assert tiling.TileSpec(core=30, halo=8).size == 46        # not a multiple of 16: allowed (resize guard path)
