"""oracle/metrics_port.py — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's `compute_metrics` (Unet-ChannalAttention.py:214-269): arg-max of the logits
(UCA:220), pixels labelled 255 dropped (UCA:223-226), TP / FP / FN per class counted with numpy (UCA:232-240) and the
four summary numbers (UCA:242-262).  Pinned against the unmodified reference function by tests/golden/metrics.npz
(oracle/make_golden.py).  Only tests/ may import it."""
from __future__ import annotations

import numpy as np
import torch


def compute_metrics(outputs: torch.Tensor, masks: torch.Tensor, num_classes: int):
    preds = torch.max(outputs, 1)[1]                                    # UCA:220 (first maximum wins)
    valid = masks != 255                                                # UCA:223
    p = preds[valid].cpu().numpy()
    m = masks[valid].cpu().numpy()
    TP, FP, FN = np.zeros(num_classes), np.zeros(num_classes), np.zeros(num_classes)
    for c in range(num_classes):                                        # UCA:236-240
        TP[c] = ((m == c) & (p == c)).sum()
        FP[c] = ((m != c) & (p == c)).sum()
        FN[c] = ((m == c) & (p != c)).sum()
    total = TP.sum() + FP.sum() + FN.sum()                              # UCA:243-245 (sic: errors count twice)
    acc = TP.sum() / total if total > 0 else 0.0
    union = TP + FP + FN
    iou = np.divide(TP, union, out=np.zeros_like(TP), where=union != 0)
    miou = np.mean(iou[union > 0]) if np.any(union > 0) else 0.0
    recall = np.divide(TP, TP + FN, out=np.zeros_like(TP), where=(TP + FN) != 0)
    mpa = np.mean(recall[(TP + FN) > 0]) if np.any((TP + FN) > 0) else 0.0
    precision = np.divide(TP, TP + FP, out=np.zeros_like(TP), where=(TP + FP) != 0)
    f1 = np.divide(2 * precision * recall, precision + recall, out=np.zeros_like(TP), where=(precision + recall) != 0)
    mf1 = np.mean(f1[(TP + FN) > 0]) if np.any((TP + FN) > 0) else 0.0
    return {"acc": float(acc), "miou": float(miou), "mpa": float(mpa), "mf1": float(mf1)}


def metric_cases():
    """Seeded (name, logits, masks, num_classes) cases shared by the golden generator and the tests."""
    g = torch.Generator().manual_seed(77)
    cases = []
    lo = torch.randn(3, 2, 24, 40, generator=g)
    ma = torch.randint(0, 2, (3, 24, 40), generator=g)
    ma[torch.rand(3, 24, 40, generator=g) < 0.05] = 255
    cases.append(("two_class", lo, ma, 2))
    lo = torch.randn(2, 4, 16, 16, generator=g)
    ma = torch.randint(0, 3, (2, 16, 16), generator=g)                  # class 3 never labelled
    ma[0, :2] = 255
    ma[1, 5, 5] = 7                                                     # a stray label value
    lo[0, :, 3, 3] = 1.5                                                # exact tie -> class 0
    cases.append(("four_class_absent_and_stray", lo, ma, 4))
    lo = torch.randn(1, 2, 8, 8, generator=g)
    cases.append(("all_ignored", lo, torch.full((1, 8, 8), 255, dtype=torch.int64), 2))
    lo = torch.zeros(1, 3, 8, 8)
    lo[:, 1] = 1.0
    cases.append(("single_class_perfect", lo, torch.ones(1, 8, 8, dtype=torch.int64), 3))
    return cases
