"""Sub-layers with the reference's names and parameter layout (T/SubLayers.py)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.init as init

from .. import ops
from .. import rng as _rng
from .Modules import AttnMask, LayerNormalization, ScaledDotProductAttention
from .Modules import BottleLinear as Linear


class MultiHeadAttention(nn.Module):
    """Per-head weight tensors w_qs/w_ks/w_vs [H, d_model, d_k] exactly as in the reference (state-dict compatible),
    but the H-fold `repeat` of q/k/v and of the mask (T/SubLayers.py:49-59: 15 % of reference decode time) is gone:
    heads are column blocks of one packed projection buffer, and the mask is a predicate."""

    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1, rng=None, site="mha"):
        super().__init__()
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        rng = rng or _rng.GLOBAL
        self.w_qs = nn.Parameter(torch.empty(n_head, d_model, d_k), requires_grad=True)
        self.w_ks = nn.Parameter(torch.empty(n_head, d_model, d_k), requires_grad=True)
        self.w_vs = nn.Parameter(torch.empty(n_head, d_model, d_v), requires_grad=True)
        self.attention = ScaledDotProductAttention(d_model, dropout, rng=rng, site=site + ".attn")
        self.layer_norm = LayerNormalization(d_model)
        self.proj = Linear(n_head * d_v, d_model)
        self.p = float(dropout)
        self._rng = rng
        self._site = rng.site(site + ".proj")
        self.return_attn = False
        self.kv_override = None      # set per call by Decoder.forward: this layer's k|v block of the fused projection
        init.xavier_normal_(self.w_qs)
        init.xavier_normal_(self.w_ks)
        init.xavier_normal_(self.w_vs)

    def forward(self, q, k, v, attn_mask: AttnMask = None):
        if attn_mask is None or attn_mask.key_pad_mask is None:
            kp = torch.ones(k.shape[0], k.shape[1], dtype=torch.uint8, device=k.device)
            attn_mask = AttnMask(kp, attn_mask.band if attn_mask is not None else None)
        # bf16 path: the residual branch of q's gradient joins the q-projection's data-gradient GEMM (ops.ResidualLink)
        link = ops.ResidualLink() if (q.dtype == torch.bfloat16 and q.requires_grad and torch.is_grad_enabled()) else None
        if q is k and k is v:
            qbuf, kvbuf = ops.head_proj(q, self.w_qs, self.w_ks, self.w_vs, link=link), None
        else:
            assert k is v, "keys and values come from the same tensor on this path (T/Layers.py:33-35)"
            kvbuf, self.kv_override = self.kv_override, None
            if kvbuf is None:
                kvbuf = ops.head_proj(k, self.w_ks, self.w_vs)
            qbuf = ops.head_proj(q, self.w_qs, link=link)
        ctx, probs = self.attention(qbuf, kvbuf, attn_mask, self.n_head, self.d_k, want_probs=self.return_attn)
        drop = self._rng.make(self.p, self._site, q.device, self.training)
        if q.dtype == torch.bfloat16:  # bf16 activation stream: tensor-core projections / attention, bf16 LayerNorm I/O
            assert q.size(1) > 1, "the bf16 path is the training path; single-token decoding runs in fp32"
            return self.layer_norm(self.proj(ctx), residual=q, drop=drop, link=link), probs
        if q.size(1) == 1:          # LayerNormalization is the identity for length-1 inputs (T/Modules.py:43-44)
            out = self.proj(ctx, drop=drop, residual=q)
        else:
            out = self.layer_norm(self.proj(ctx), residual=q, drop=drop)
        if probs is not None:       # reference layout: [n_head * batch, len_q, len_k], head-major
            probs = probs.permute(1, 0, 2, 3).reshape(-1, probs.shape[2], probs.shape[3])
        return out, probs


class PositionwiseFeedForward(nn.Module):
    """LN(dropout(W2 relu(W1 x + b1) + b2) + x); the two Conv1d(k=1) of T/SubLayers.py:70-86 are position-wise linear
    maps, so their [out, in, 1] weights feed the GEMM kernel directly (no transposes)."""

    def __init__(self, d_hid, d_inner_hid, dropout=0.1, rng=None, site="ffn"):
        super().__init__()
        self.w_1 = nn.Conv1d(d_hid, d_inner_hid, 1)
        self.w_2 = nn.Conv1d(d_inner_hid, d_hid, 1)
        self.layer_norm = LayerNormalization(d_hid)
        self.p = float(dropout)
        self._rng = rng or _rng.GLOBAL
        self._site = self._rng.site(site)

    def forward(self, x):
        drop = self._rng.make(self.p, self._site, x.device, self.training)
        if x.dtype == torch.bfloat16:  # both GEMMs on tcgen05 with fused bias(+ReLU) epilogues
            link = ops.ResidualLink() if (x.requires_grad and torch.is_grad_enabled()) else None
            h = ops.linear_tc(x, self.w_1.weight, self.w_1.bias, relu=True, link=link)
            y = ops.linear_tc(h, self.w_2.weight, self.w_2.bias)
            return self.layer_norm(y, residual=x, drop=drop, link=link)
        h = ops.linear(x, self.w_1.weight, self.w_1.bias, relu=True)
        if x.size(1) == 1:
            return ops.linear(h, self.w_2.weight, self.w_2.bias, drop=drop, residual=x)
        y = ops.linear(h, self.w_2.weight, self.w_2.bias)
        return self.layer_norm(y, residual=x, drop=drop)
