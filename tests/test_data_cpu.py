"""Input pipeline: the drop-in VOCSegDataset returns what the reference's class returns
(tests/golden/voc_mini_expected.npz, produced by oracle/make_golden.py from the unmodified reference), and the raw
uint8 mode + the device preprocessing formula reproduce the same tensors bit for bit."""
import os

import numpy as np
import pytest
import torch
import torchvision.transforms as T

from unetca_b200 import data


@pytest.fixture
def voc(golden_dir):
    return os.path.join(golden_dir, "voc_mini"), np.load(os.path.join(golden_dir, "voc_mini_expected.npz"))


def test_dataset_matches_reference_items(voc):
    root, g = voc
    tf = T.Compose([T.Resize((32, 32)), T.ToTensor(), T.Normalize(mean=[0.5], std=[0.5])])
    for split, n in (("train", 4), ("val", 2)):
        ds = data.VOCSegDataset(root, 32, image_set=split, transforms=tf)
        assert len(ds) == n
        for i, name in enumerate(ds.ids):
            img, mask = ds[i]
            assert img.dtype == torch.float32 and mask.dtype == torch.int64
            assert np.array_equal(img.numpy(), g[f"img:{name}"]), name             # bit-exact
            assert np.array_equal(mask.numpy(), g[f"mask:{name}"]), name
            assert set(np.unique(mask.numpy())) <= {0, 1}                         # 255 -> 1, anything else -> 0


def test_raw_mode_plus_prep_formula_is_bit_exact(voc):
    """raw=True stops before ToTensor; ((u8 / 255) - 0.5) / 0.5 and (long)(u8 / 255) — the arithmetic of
    unetca_prep_u8 — give the reference tensors exactly."""
    root, g = voc
    ds = data.VOCSegDataset(root, 32, image_set="train", raw=True)
    for i, name in enumerate(ds.ids):
        img, mask = ds[i]
        assert img.dtype == torch.uint8 and mask.dtype == torch.uint8 and img.shape == (32, 32)
        x = ((img.to(torch.float32) / 255.0) - 0.5) / 0.5
        y = (mask.to(torch.float32) / 255.0).long()
        assert np.array_equal(x[None].numpy(), g[f"img:{name}"]), name
        assert np.array_equal(y.numpy(), g[f"mask:{name}"]), name


def test_missing_image_set_raises(voc):
    root, _ = voc
    with pytest.raises(FileNotFoundError):
        data.VOCSegDataset(root, 32, image_set="test")
