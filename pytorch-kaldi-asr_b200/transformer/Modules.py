"""Leaf modules with the reference's names (T/Modules.py), each a thin shell over one fused sm_100a kernel."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.init as init

from .. import ops
from .. import rng as _rng


class Linear(nn.Module):
    """nn.Linear holder with xavier-normal init (T/Modules.py:8-17); accepts any number of leading dimensions, so the
    reference's Bottle reshape dance (T/Modules.py:18-30, the `.view` that breaks on torch>=2) is unnecessary."""

    def __init__(self, d_in, d_out, bias=True):
        super().__init__()
        self.linear = nn.Linear(d_in, d_out, bias=bias)
        init.xavier_normal_(self.linear.weight)

    def forward(self, x, drop=None, residual=None, out_fp32=False, single_consumer=False):
        if x.dtype == torch.bfloat16:          # bf16 training path: tcgen05 GEMMs (ops.set_compute_mode("bf16"))
            assert residual is None
            return ops.linear_tc(x, self.linear.weight, self.linear.bias, drop=drop, out_fp32=out_fp32,
                                 single_consumer=single_consumer)
        return ops.linear(x, self.linear.weight, self.linear.bias, drop=drop, residual=residual)


BottleLinear = Linear


class LayerNormalization(nn.Module):
    """(z-mu)/(sigma_unbiased+eps)*a_2+b_2 with eps=1e-3; identity when z.size(1)==1 (T/Modules.py:32-51)."""

    def __init__(self, d_hid, eps=1e-3):
        super().__init__()
        self.eps = eps
        self.a_2 = nn.Parameter(torch.ones(d_hid), requires_grad=True)
        self.b_2 = nn.Parameter(torch.zeros(d_hid), requires_grad=True)

    def forward(self, z, residual=None, drop=None, link=None):
        """LN(dropout(z) + residual); the extras default to the plain reference call LN(z).  Length-1 inputs (where
        the reference skips normalisation) are handled by the callers, which fold the residual into their GEMM.
        `link` (ops.ResidualLink): the residual branch of the gradient is handed to the sub-layer's first GEMM."""
        if z.size(1) == 1:
            assert residual is None and drop is None, "length-1 residual path is fused into the producing GEMM"
            return z
        return ops.add_layer_norm(z, residual, self.a_2, self.b_2, self.eps, drop, link)


class AttnMask:
    """Symbolic attention mask: key padding (uint8 [B,Lk], 1 = real) and/or a band (start,end) meaning query i may see
    keys i+start..i+end.  It stands in for the dense [B,Lq,Lk] tensors of T/Models.py:27-49, which the kernels evaluate
    as a predicate instead of reading from HBM.  `a + b` merges two masks (the reference ORs them via torch.gt(a+b,0))."""

    def __init__(self, key_pad_mask=None, band=None, q_len=None):
        self.key_pad_mask, self.band, self.q_len = key_pad_mask, band, q_len

    def __add__(self, other):
        band = self.band
        if other.band is not None:
            band = other.band if band is None else (max(band[0], other.band[0]), min(band[1], other.band[1]))
        return AttnMask(self.key_pad_mask if self.key_pad_mask is not None else other.key_pad_mask, band,
                        self.q_len or other.q_len)

    def dense(self):
        """Materialise the reference's boolean mask (True = masked) -- debugging / tests only."""
        kp = self.key_pad_mask
        b, lk = kp.shape
        lq = self.q_len or lk
        masked = kp.eq(0)[:, None, :].expand(b, lq, lk).clone()
        if self.band is not None:
            i = torch.arange(lq, device=kp.device)[:, None]
            j = torch.arange(lk, device=kp.device)[None, :]
            masked |= ((j < i + self.band[0]) | (j > i + self.band[1]))[None]
        return masked


class ScaledDotProductAttention(nn.Module):
    """softmax(mask(q k^T / sqrt(d_model))) v with dropout on the probabilities (T/Modules.py:67-97), as ONE flash-style
    kernel over packed projections.  Note the scale is 1/sqrt(d_model), not 1/sqrt(d_k) (T/Modules.py:72)."""

    def __init__(self, d_model, attn_dropout=0.1, rng=None, site="attn"):
        super().__init__()
        self.temper = math.sqrt(d_model)
        self.p = float(attn_dropout)
        self._rng = rng or _rng.GLOBAL
        self._site = self._rng.site(site)

    def forward(self, qbuf, kvbuf, mask: AttnMask, n_head, d_k, want_probs=False):
        drop = self._rng.make(self.p, self._site, qbuf.device, self.training)
        if qbuf.dtype == torch.bfloat16:
            assert not want_probs, "the attention maps are only materialised on the fp32 path"
            ctx, _lse = ops.attention_tc(qbuf, kvbuf, mask.key_pad_mask, n_head, d_k, mask.band, 1.0 / self.temper, drop)
            return ctx, None
        return ops.attention(qbuf, kvbuf, mask.key_pad_mask, n_head, d_k, mask.band, 1.0 / self.temper, drop, want_probs)
