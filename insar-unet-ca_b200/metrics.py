"""On-device `compute_metrics` of the reference (Unet-ChannalAttention.py:214-269).

The reference takes the arg-max of the logits, drops pixels labelled 255, copies BOTH masks to the host and counts
TP / FP / FN per class with numpy (UCA:220-240) — a device-to-host copy and a numpy pass that serialise every
validation batch.  Here the counting runs on the GPU (`unetca_confusion_counts`: a (num_classes+1) x num_classes
label-by-prediction table); only that table crosses to the host, and the four numbers are formed from it with the
reference's own formulas, including its conventions: `acc` divides by TP+FP+FN (a misclassified pixel is counted
twice, UCA:243-245), classes without pixels are left out of the means, label values other than 0..num_classes-1 and
255 still count as false positives of the predicted class.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import _lib


def confusion_counts(outputs: torch.Tensor, masks: torch.Tensor, num_classes: int, ignore_index: int = 255) -> torch.Tensor:
    """(num_classes+1, num_classes) int64 CUDA tensor: row = label (last row: other label values), column = arg-max."""
    if not outputs.is_cuda or not masks.is_cuda:
        raise RuntimeError("unetca_b200.metrics runs only on CUDA tensors; there is no CPU fallback")
    if outputs.dim() != 4 or outputs.shape[1] != num_classes:
        raise ValueError(f"outputs must be (B, {num_classes}, H, W)")
    B, nc, H, W = outputs.shape
    if masks.shape != (B, H, W) or masks.dtype != torch.int64:
        raise ValueError(f"masks must be an int64 tensor of shape {(B, H, W)}")
    logits = outputs.detach().float().contiguous()
    target = masks.detach().contiguous()
    lib = _lib.load()
    ncell = (nc + 1) * nc
    parts = torch.empty(lib.unetca_max_parts(B) * ncell, dtype=torch.int64, device=logits.device)
    counts = torch.empty(nc + 1, nc, dtype=torch.int64, device=logits.device)
    _lib.call("unetca_confusion_counts", logits.data_ptr(), target.data_ptr(), nc, B, H * W, int(ignore_index),
              parts.data_ptr(), counts.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return counts


def metrics_from_counts(counts: np.ndarray) -> Dict[str, float]:
    """The reference's formulas (UCA:232-269) on a (num_classes+1, num_classes) label-by-prediction table."""
    counts = np.asarray(counts, dtype=np.float64)
    nc = counts.shape[1]
    TP = np.array([counts[c, c] for c in range(nc)])
    FP = counts.sum(axis=0) - TP                         # predicted c, label != c (UCA:238)
    FN = counts[:nc].sum(axis=1) - TP                    # label c, predicted != c (UCA:239)
    total = TP.sum() + FP.sum() + FN.sum()
    acc = TP.sum() / total if total > 0 else 0.0
    union = TP + FP + FN
    iou = np.divide(TP, union, out=np.zeros_like(TP), where=union != 0)
    miou = float(np.mean(iou[union > 0])) if np.any(union > 0) else 0.0
    recall = np.divide(TP, TP + FN, out=np.zeros_like(TP), where=(TP + FN) != 0)
    mpa = float(np.mean(recall[(TP + FN) > 0])) if np.any((TP + FN) > 0) else 0.0
    precision = np.divide(TP, TP + FP, out=np.zeros_like(TP), where=(TP + FP) != 0)
    f1 = np.divide(2 * precision * recall, precision + recall, out=np.zeros_like(TP), where=(precision + recall) != 0)
    mf1 = float(np.mean(f1[(TP + FN) > 0])) if np.any((TP + FN) > 0) else 0.0
    return {"acc": float(acc), "miou": miou, "mpa": mpa, "mf1": mf1}


def compute_metrics(outputs: torch.Tensor, masks: torch.Tensor, num_classes: int) -> Dict[str, float]:
    """Drop-in for the reference's `compute_metrics(outputs, masks, num_classes)` (UCA:214): same dict, same values."""
    return metrics_from_counts(confusion_counts(outputs, masks, num_classes).cpu().numpy())
