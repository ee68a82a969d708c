"""CPU restatement of the reference's training loss and optimiser step (row f.2 of SURVEY.md §8).

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's CPU legs and oracle/gen_golden.py; never by the
product path (segtran_b200/).  Pinned against the real reference (utils/losses.py, optimization.py run from
/root/reference in the build container) by tests/golden/train_loss_tiny.pt and train_bertadam_tiny.pt.

  seg_loss       train3d.py:731-756 (BCEWithLogits(pos_weight) on [B,*,K] + per-class dice_loss_indiv on the
                 sigmoid, utils/losses.py:47-60), with the loss weights of train3d.py:685-697, :515-518
  clip_grad_norm train3d.py:760-761 (torch.nn.utils.clip_grad_norm_ over all parameters)
  bert_adam_step optimization.py:90-164 (per-parameter norm clip, no bias correction, decoupled decay,
                 warm-up schedules :11-37)
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------------
# loss
# ---------------------------------------------------------------------------------------------------------
def default_class_weights(num_classes: int, focus_class: int = -1) -> torch.Tensor:
    """train3d.py:686-690: background 0, the others equal (focus class doubled), normalised to sum 1."""
    w = torch.ones(num_classes)
    w[0] = 0
    if focus_class != -1:
        w[focus_class] = 2
    return w / w.sum()


def normalised_bce_weight(bce_weight: Sequence[float], num_classes: int) -> torch.Tensor:
    """train3d.py:517-518: pos_weight = w * (K-1) / sum(w)."""
    w = torch.tensor(list(bce_weight), dtype=torch.float32)
    return w * (num_classes - 1) / w.sum()


def dice_loss_indiv(score: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """utils/losses.py:47-60: per-sample soft Dice with squared denominators, smooth 1e-5, mean over the batch."""
    score = score.reshape(score.shape[0], -1)
    gt = gt.float().reshape(gt.shape[0], -1)
    smooth = 1e-5
    inter = (score * gt).sum(1)
    dice = (2 * inter + smooth) / ((score * score).sum(1) + (gt * gt).sum(1) + smooth)
    return (1 - dice).mean()


def seg_loss(logits: torch.Tensor, mask: torch.Tensor, pos_weight: Optional[torch.Tensor],
             class_weights: torch.Tensor, dice_w: float = 0.5):
    """logits, mask: [B,K,*spatial] (mask n-hot float).  Returns (loss, ce, dice_total)   (train3d.py:738-756)."""
    K = logits.shape[1]
    perm = [0] + list(range(2, logits.dim())) + [1]                       # class dim last for pos_weight (:738-742)
    ce = F.binary_cross_entropy_with_logits(logits.permute(perm), mask.permute(perm), pos_weight=pos_weight)
    soft = torch.sigmoid(logits)
    dice_total = logits.new_zeros(())
    for cls in range(1, K):                                               # :746-751 (class 0 = background skipped)
        dice_total = dice_total + dice_loss_indiv(soft[:, cls], mask[:, cls]) * class_weights[cls]
    loss = (1 - dice_w) * ce + dice_w * dice_total                        # :756 (attention-consistency term off)
    return loss, ce, dice_total


# ---------------------------------------------------------------------------------------------------------
# optimiser
# ---------------------------------------------------------------------------------------------------------
def warmup_linear(x: float, warmup: float) -> float:                     # optimization.py:25-31
    if x < warmup:
        return x / warmup
    return max((x - 1.0) / (warmup - 1.0), 0.0)


def warmup_constant(x: float, warmup: float) -> float:                   # optimization.py:16-23
    return x / warmup if x < warmup else 1.0


SCHEDULES = {"warmup_linear": warmup_linear, "warmup_constant": warmup_constant}


def clip_grad_norm(grads: List[torch.Tensor], max_norm: float) -> float:
    """torch.nn.utils.clip_grad_norm_ (train3d.py:760-761): one global L2 norm, grads scaled in place."""
    total = math.sqrt(sum(float(g.double().pow(2).sum()) for g in grads))
    coef = min(max_norm / (total + 1e-6), 1.0)
    for g in grads:
        g.mul_(coef)
    return total


def bert_adam_step(params: List[torch.Tensor], grads: List[torch.Tensor], state: Dict, *, lr: Sequence[float],
                   weight_decay: Sequence[float], warmup: float = -1, t_total: int = -1,
                   schedule: str = "warmup_linear", b1: float = 0.9, b2: float = 0.999, e: float = 1e-6,
                   max_grad_norm: float = 0.05):
    """One BertAdam.step() over a flat list of parameters (optimization.py:90-164).  `lr` / `weight_decay` are
    per-parameter (the reference's param groups, train3d.py:334-339).  `state` carries step, m, v."""
    if not state:
        state["step"] = 0
        state["m"] = [torch.zeros_like(p) for p in params]
        state["v"] = [torch.zeros_like(p) for p in params]
    for i, (p, g) in enumerate(zip(params, grads)):
        if g is None:
            continue
        if max_grad_norm > 0:                                             # :120-121 per-parameter clip, in place
            norm = float(g.norm(2))
            coef = min(max_grad_norm / (norm + 1e-6), 1.0)
            g.mul_(coef)
        m, v = state["m"][i], state["v"][i]
        m.mul_(b1).add_(g, alpha=1 - b1)                                  # :125
        v.mul_(b2).addcmul_(g, g, value=1 - b2)                           # :126
        update = m / (v.sqrt() + e)                                       # :127 (no bias correction, :159-162)
        if weight_decay[i] > 0.0:
            update = update + weight_decay[i] * p                         # :136-137 decoupled decay
        if t_total != -1:
            lr_s = lr[i] * SCHEDULES[schedule](state["step"] / t_total, warmup)     # :139-142
        else:
            lr_s = lr[i]
        p.add_(-lr_s * update)                                            # :154-155
    state["step"] += 1                                                    # :157 (per-parameter counters move together)
