"""Checkpoints: the reference's best-model file plus a resumable training state.

The reference only ever writes `torch.save(model.state_dict(), MODEL_SAVE_PATH)` when the validation mIoU improves
(Unet-ChannalAttention.py:382-387) and cannot resume: optimizer moments, the epoch counter, the best mIoU and the
history list are lost with the process.  `save_best` writes exactly the reference's file (same 154 / 136 keys, OIHW
fp32 — it loads into the reference's own UNet and vice versa); `save_resume` / `load_resume` add what a long
data-parallel run needs to restart.  Files are written to a temporary name and renamed, so a killed job never leaves a
truncated checkpoint behind.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Tuple

import torch

FORMAT = "unetca_b200.resume.v1"


def _atomic_save(obj, path: str) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    tmp = f"{path}.tmp.{os.getpid()}"
    torch.save(obj, tmp)
    os.replace(tmp, path)


def save_best(model: torch.nn.Module, path: str) -> None:
    """The reference's checkpoint (UCA:386): the bare state_dict."""
    _atomic_save(model.state_dict(), path)


def save_resume(path: str, model: torch.nn.Module, optimizer: torch.optim.Optimizer, epoch: int,
                best_m_iou: float = -1.0, history: Optional[List[Dict[str, Any]]] = None) -> None:
    """Everything `train_model` (UCA:320-399) keeps in local variables, after `epoch` completed epochs."""
    _atomic_save({"format": FORMAT, "model": model.state_dict(), "optimizer": optimizer.state_dict(), "epoch": int(epoch),
                  "best_m_iou": float(best_m_iou), "history": list(history or []),
                  "precision": getattr(model, "precision", None)}, path)


def load_resume(path: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None,
                map_location=None, trust_pickle: bool = False) -> Tuple[int, float, List[Dict[str, Any]]]:
    """Restore a `save_resume` file (or a bare reference state_dict: model only).  Returns (completed epochs,
    best mIoU so far, history) — resume the loop at `range(epoch, num_epochs)` with `best_m_iou` as in UCA:326.

    Files are read with torch's safe unpickler (`weights_only=True`): everything `save_resume` writes is tensors and
    plain Python types.  Only `trust_pickle=True` falls back to full pickle (arbitrary code execution — for files whose
    history holds custom objects and whose origin is trusted)."""
    try:
        obj = torch.load(path, map_location=map_location, weights_only=True)
    except Exception:
        if not trust_pickle:
            raise
        obj = torch.load(path, map_location=map_location, weights_only=False)
    if isinstance(obj, dict) and obj.get("format") == FORMAT:
        model.load_state_dict(obj["model"])
        if optimizer is not None:
            optimizer.load_state_dict(obj["optimizer"])
        if obj.get("precision") and hasattr(model, "set_precision"):
            model.set_precision(obj["precision"])
        return obj["epoch"], obj["best_m_iou"], obj["history"]
    model.load_state_dict(obj)              # a reference checkpoint (UCA:386)
    return 0, -1.0, []
