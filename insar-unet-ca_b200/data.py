"""Input pipeline for the reference's tile datasets (Unet-ChannalAttention.py:165-212, 428-451).

`VOCSegDataset` is a drop-in for the reference's class (same constructor, same files, same `(image, mask)` items when a
`transforms` pipeline is passed).  With `raw=True` it stops before `ToTensor`: items are uint8 arrays (image resized
with PIL bilinear exactly as `T.Resize` does on a PIL image, mask resized NEAREST), which is all the host has to do.
`DevicePrefetcher` then moves uint8 batches (2 bytes per pixel instead of the 12 of fp32 image + int64 mask) from
pinned memory on a side stream, double-buffered, and finishes the reference's preprocessing on the GPU in one kernel
(`unetca_prep_u8`): `ToTensor` (/255), `Normalize(mean, std)` and the mask's `ToTensor().long()` — which maps 255 to 1
and everything else to 0 (UCA:208-210).  Bit-identical to the reference's tensors (same IEEE operations).
"""
from __future__ import annotations

import os
from typing import Iterable, Iterator, Optional, Tuple

import numpy as np
import torch
from PIL import Image
from torch.utils.data import Dataset

from . import _lib


class VOCSegDataset(Dataset):
    """VOC-style tile dataset of the reference (UCA:166-212): JPEGImages/<id>.jpg, SegmentationClass/<id>.png,
    ImageSets/Segmentation/<image_set>.txt."""

    def __init__(self, voc_root: str, image_size: int, image_set: str = "train", transforms=None, raw: bool = False):
        self.voc_root, self.image_size, self.transforms, self.raw = voc_root, image_size, transforms, raw
        self.image_dir = os.path.join(voc_root, "JPEGImages")
        self.mask_dir = os.path.join(voc_root, "SegmentationClass")
        self.image_set_path = os.path.join(voc_root, "ImageSets", "Segmentation", f"{image_set}.txt")
        if not os.path.exists(self.image_set_path):
            raise FileNotFoundError(f"ImageSets file not found: {self.image_set_path}")
        with open(self.image_set_path) as f:
            self.ids = [line.strip() for line in f.readlines()]

    def __len__(self) -> int:
        return len(self.ids)

    def _load(self, idx: int):
        img_id = self.ids[idx]
        img = Image.open(os.path.join(self.image_dir, f"{img_id}.jpg")).convert("L")          # UCA:195
        mask = Image.open(os.path.join(self.mask_dir, f"{img_id}.png")).convert("L")          # UCA:198
        return img, mask

    def __getitem__(self, idx: int):
        img, mask = self._load(idx)
        size = (self.image_size, self.image_size)
        mask = mask.resize(size[::-1], Image.NEAREST)                                          # UCA:204-206
        if self.raw:
            # T.Resize((S, S)) on a PIL image is PIL's BILINEAR resize (reducing filter included)
            img = img.resize(size[::-1], Image.BILINEAR)
            return torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()), torch.from_numpy(np.asarray(mask, dtype=np.uint8).copy())
        if self.transforms is not None:
            img = self.transforms(img)                                                         # UCA:200-201
        m = torch.from_numpy(np.asarray(mask, dtype=np.uint8).copy())
        mask_t = (m.to(torch.float32).div(255)).long()                                         # ToTensor().long(), UCA:208-210
        return img, mask_t


def prep_u8(images_u8: torch.Tensor, masks_u8: torch.Tensor, mean: float = 0.5, std: float = 0.5) -> Tuple[torch.Tensor, torch.Tensor]:
    """uint8 CUDA batches (B,H,W) -> (fp32 (B,1,H,W) normalised images, int64 (B,H,W) masks) as the reference's
    `T.ToTensor(); T.Normalize([mean],[std])` and `T.ToTensor()(mask).squeeze(0).long()` produce them."""
    if not images_u8.is_cuda or not masks_u8.is_cuda:
        raise RuntimeError("unetca_b200.data.prep_u8 runs only on CUDA tensors; there is no CPU fallback")
    if (images_u8.dtype != torch.uint8 or masks_u8.dtype != torch.uint8 or images_u8.dim() != 3
            or images_u8.shape != masks_u8.shape):
        raise ValueError("images and masks must be uint8 tensors of the same (B, H, W) shape")
    images_u8, masks_u8 = images_u8.contiguous(), masks_u8.contiguous()
    B, H, W = images_u8.shape
    out = torch.empty(B, 1, H, W, dtype=torch.float32, device=images_u8.device)
    lab = torch.empty(B, H, W, dtype=torch.int64, device=images_u8.device)
    _lib.call("unetca_prep_u8", images_u8.data_ptr(), masks_u8.data_ptr(), out.data_ptr(), lab.data_ptr(), B * H * W,
              float(mean), float(std), torch.cuda.current_stream().cuda_stream)
    return out, lab


class DevicePrefetcher:
    """Wraps an iterable of uint8 `(images (B,H,W), masks (B,H,W))` CPU batches (e.g. a DataLoader over
    `VOCSegDataset(raw=True)`): pinned staging buffers, H2D copies on a side stream one batch ahead, device-side
    preprocessing.  Yields `(images fp32 (B,1,H,W), masks int64 (B,H,W))` on `device`."""

    def __init__(self, loader: Iterable, device: torch.device, mean: float = 0.5, std: float = 0.5):
        self.loader, self.device, self.mean, self.std = loader, torch.device(device), mean, std
        self.stream = torch.cuda.Stream(device=self.device)

    def _stage(self, batch):
        img, mask = batch
        img = img if img.is_pinned() else img.pin_memory()
        mask = mask if mask.is_pinned() else mask.pin_memory()
        with torch.cuda.stream(self.stream):
            dimg = img.to(self.device, non_blocking=True)
            dmask = mask.to(self.device, non_blocking=True)
            out = prep_u8(dimg, dmask, self.mean, self.std)
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return out, ev, (img, mask)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        it = iter(self.loader)
        nxt: Optional[tuple] = None
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            (x, y), ev, _keep = nxt
            try:
                nxt = self._stage(next(it))          # the next batch's copy overlaps this batch's step
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(ev)
            x.record_stream(torch.cuda.current_stream(self.device))
            y.record_stream(torch.cuda.current_stream(self.device))
            yield x, y
