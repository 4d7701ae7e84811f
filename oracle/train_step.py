"""Restatement of one reference training step (L/train.py:145-207) and its optimiser (T/Optim.py:4-27 around
torch.optim.Adam, L/train.py:376-380).  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import acoustic_model as am


class AdamSchedule:
    """Adam(beta=(0.9,0.999), eps=1e-8, no weight decay) written out by hand over a name->tensor dict, plus the
    reference's LR rule: lr_n = start_lr*c/(n+c) is installed *after* step n; step 1 runs at Adam's constructor
    default 1e-3 (T/Optim.py:21-27, L/train.py:376-380)."""

    def __init__(self, params: Dict[str, torch.Tensor], start_lr: float = 1e-3, soft_coefficient: float = 500,
                 betas=(0.9, 0.999), eps: float = 1e-8):
        self.params = params
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.betas, self.eps = betas, eps
        self.start_lr, self.c = start_lr, soft_coefficient
        self.lr = 1e-3
        self.t = 0                      # Adam's own step counter
        self.n_current_steps = 0

    @torch.no_grad()
    def step(self, grads: Dict[str, torch.Tensor]):
        b1, b2 = self.betas
        self.t += 1
        bc1 = 1.0 - b1 ** self.t
        bc2 = 1.0 - b2 ** self.t
        for k, p in self.params.items():
            g = grads.get(k)
            if g is None:
                continue
            self.m[k].mul_(b1).add_(g, alpha=1 - b1)
            self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (self.v[k].sqrt() / (bc2 ** 0.5)).add_(self.eps)
            p.addcdiv_(self.m[k], denom, value=-self.lr / bc1)

    def update_learning_rate(self):
        self.n_current_steps += 1
        self.lr = (self.start_lr * self.c) / (self.n_current_steps + self.c)


def loss_and_grads(sd, cfg, batch, smoothing: bool = False, drop: Optional[am.DropoutPlan] = None):
    """Forward + summed CE + autograd backward.  `batch` = (src f32[B,T,F], src_mask u8[B,T], tgt i64[B,L+1],
    tgt_mask u8[B,L+1]); teacher forcing split per L/train.py:163-165."""
    src, src_mask, tgt, tgt_mask = [torch.as_tensor(x) for x in batch]
    goal, tgt_in, tgt_in_mask = tgt[:, 1:], tgt[:, :-1], tgt_mask[:, :-1]
    keys = am.trainable_keys(sd)
    leaf = {k: (sd[k].detach().clone().requires_grad_(True) if k in keys else sd[k]) for k in sd}
    logits = am.transformer_forward(leaf, cfg, src.float(), src_mask, tgt_in.long(), tgt_in_mask, drop)
    loss, n_correct, n_words = am.performance(logits, goal.long(), smoothing)
    loss.backward()
    grads = {k: leaf[k].grad for k in keys}
    return logits.detach(), loss.detach(), int(n_correct), int(n_words), grads


def train_steps(sd, cfg, batches, start_lr=1e-3, soft_coefficient=25000, smoothing=False, drop_mode="off"):
    """Run len(batches) optimiser steps in place on `sd`; returns the per-step summed losses."""
    keys = am.trainable_keys(sd)
    opt = AdamSchedule({k: sd[k] for k in keys}, start_lr, soft_coefficient)
    losses = []
    for batch in batches:
        drop = am.DropoutPlan(drop_mode)
        _, loss, _, _, grads = loss_and_grads(sd, cfg, batch, smoothing, drop)
        opt.step(grads)
        opt.update_learning_rate()
        losses.append(float(loss))
    return losses
