"""Dense per-point losses of the GAPartNet step, restated with plain torch
(/root/reference/gapartnet/network/losses.py: focal_loss :35-64, dice_loss :132-158 (kornia-style),
pixel_accuracy :8-20).  Not part of the sparse-conv hot path but needed to
assemble the full train step (BASELINE config #4)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F


@torch.no_grad()
def pixel_accuracy(pred: torch.Tensor, gt: torch.Tensor) -> float:
    return float((pred == gt).sum() / gt.numel()) if gt.numel() > 0 else 0.0


def focal_loss(inputs: torch.Tensor, targets: torch.Tensor, alpha: Optional[torch.Tensor] = None, gamma: float = 2.0,
               reduction: str = "mean", ignore_index: int = -100) -> torch.Tensor:
    keep = targets != ignore_index
    targets = targets[keep]
    if targets.shape[0] == 0:
        return inputs.new_zeros(())
    log_p = F.log_softmax(inputs[keep], dim=-1)
    ce = F.nll_loss(log_p, targets, weight=alpha, ignore_index=ignore_index, reduction="none")
    log_pt = log_p.gather(1, targets[:, None]).squeeze(-1)
    loss = ce * (1 - log_pt.exp()) ** gamma
    return loss.mean() if reduction == "mean" else (loss.sum() if reduction == "sum" else loss)


def dice_loss(logits: torch.Tensor, target: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """logits [B,C,H,W], target [B,H,W] int64; soft dice over (C,H,W) per batch item with the
    reference's one-hot smoothing (+1e-6)"""
    soft = F.softmax(logits, dim=1)
    onehot = torch.zeros_like(soft).scatter_(1, target.unsqueeze(1), 1.0) + 1e-6
    inter = (soft * onehot).sum((1, 2, 3))
    card = (soft + onehot).sum((1, 2, 3))
    return (1.0 - 2.0 * inter / (card + eps)).mean()
