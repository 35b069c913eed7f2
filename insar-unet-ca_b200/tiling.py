"""Large-scene sliding-window inference, sharded by spatial tiles with halo (BASELINE.json configs[3]).

The reference has no large-scene path (scenes are pre-cut into 64/128-px tiles outside the repo, UCA:17,169); its
definition of "the model's output on a tile" is `model.eval(); model(tile)` (validate_model, UCA:273-287).  Because
the SE squeeze (UCA:65) averages over the *whole input*, a tile's result depends on the tile's own extent, so the
output of tiled inference is defined as: reference eval-forward on (core + halo), zero-filled outside the scene, keep
the core's argmax (UCA:220).  The parity tests check exactly that definition, tile for tile.

Tiles are independent: rank r of `world` takes tiles r, r+world, ... (static round-robin), no collective on the data
path; the optional gather of the uint8 mask is left to the caller.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Iterator, List, Tuple

import torch


@dataclass(frozen=True)
class TileSpec:
    core: int = 1024          # pixels of output kept per tile (square)
    halo: int = 128           # context on each side; the conv stack's receptive field is 200 px -> halo >= 100

    def __post_init__(self):
        if self.core <= 0 or self.halo < 0 or self.core + 2 * self.halo < 16:
            raise ValueError("core + 2*halo must be at least 16 (four 2x2 poolings); multiples of 16 avoid the "
                             "decoder's resize guard (UCA:138-157) and are the fast path")

    @property
    def size(self) -> int:
        return self.core + 2 * self.halo


@dataclass(frozen=True)
class Tile:
    index: int
    y0: int                   # top-left of the core in scene coordinates
    x0: int
    h: int                    # core extent actually inside the scene
    w: int


def plan(H: int, W: int, spec: TileSpec) -> List[Tile]:
    """Row-major list of tiles whose cores partition the H x W scene."""
    tiles, idx = [], 0
    for y0 in range(0, H, spec.core):
        for x0 in range(0, W, spec.core):
            tiles.append(Tile(idx, y0, x0, min(spec.core, H - y0), min(spec.core, W - x0)))
            idx += 1
    return tiles


def shard(tiles: List[Tile], rank: int, world: int) -> List[Tile]:
    return tiles[rank::world]


def extract(scene: torch.Tensor, tile: Tile, spec: TileSpec) -> torch.Tensor:
    """(C, size, size) window around the tile's core, zero-filled outside the (C,H,W) scene."""
    C, H, W = scene.shape
    out = torch.zeros(C, spec.size, spec.size, dtype=scene.dtype, device=scene.device)
    ys, xs = tile.y0 - spec.halo, tile.x0 - spec.halo
    y_lo, x_lo = max(ys, 0), max(xs, 0)
    y_hi, x_hi = min(ys + spec.size, H), min(xs + spec.size, W)
    out[:, y_lo - ys:y_hi - ys, x_lo - xs:x_hi - xs] = scene[:, y_lo:y_hi, x_lo:x_hi]
    return out


@torch.no_grad()
def predict_tiles(model, tiles: List[Tile], spec: TileSpec, fetch: Callable[[Tile], torch.Tensor],
                  batch: int = 1) -> Iterator[Tuple[Tile, torch.Tensor]]:
    """Yield (tile, uint8 class map of the tile's core) for every tile; `fetch(tile)` returns its (C,size,size)
    window on the model's device.  Runs the eval-mode CUDA path (`model.predict_mask`), `batch` tiles at a time."""
    was_training = model.training
    model.eval()
    try:
        for i in range(0, len(tiles), batch):
            group = tiles[i:i + batch]
            x = torch.stack([fetch(t) for t in group])
            mask = model.predict_mask(x)
            for t, m in zip(group, mask):
                yield t, m[spec.halo:spec.halo + t.h, spec.halo:spec.halo + t.w].to(torch.uint8)
    finally:
        model.train(was_training)


@torch.no_grad()
def predict_scene(model, scene: torch.Tensor, spec: TileSpec, rank: int = 0, world: int = 1, batch: int = 1,
                  out: torch.Tensor | None = None) -> torch.Tensor:
    """Class map (H, W) uint8 of a (C,H,W) scene resident on the model's device; with world > 1 only this rank's
    tiles are filled (the rest stay 255)."""
    C, H, W = scene.shape
    if out is None:
        out = torch.full((H, W), 255, dtype=torch.uint8, device=scene.device)
    tiles = shard(plan(H, W, spec), rank, world)
    for t, m in predict_tiles(model, tiles, spec, lambda tl: extract(scene, tl, spec), batch):
        out[t.y0:t.y0 + t.h, t.x0:t.x0 + t.w] = m
    return out
