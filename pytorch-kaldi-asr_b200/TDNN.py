"""Front-end and TDNN building blocks with the reference's class names (L/pytorch/TDNN.py), backed by sm_100a kernels.

ConcatLayer never materialises the spliced tensor when it feeds a TDNNLayer: the splice becomes a per-context row
shift inside the GEMM's tile loader (fp32 path) / a 3-D TMA box with out-of-bounds zero fill (bf16 tcgen05 path).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.init as init

from . import ops
from . import rng as _rng


class ConcatLayer(nn.Module):
    """Frame splicing with zero padding at the tensor edges (L/pytorch/TDNN.py:6-28).  Stand-alone it runs the fused
    front-end kernel (fold=1, no CMVN); inside TDNNLayer / EncoderTest it is folded into the consumer."""

    def __init__(self, index=(0,)):
        super().__init__()
        self.index = [int(i) for i in index]
        self.pad_head = max(0, -self.index[0])
        self.pad_end = max(0, self.index[-1])

    def forward(self, x):
        return ops.frontend(x, None, 1, self.index, 0)


class TDNNLayer(nn.Module):
    """splice(ctx) -> Linear(+bias) -> ReLU -> dropout, one fused GEMM launch (L/pytorch/TDNN.py:31-46)."""

    def __init__(self, d_input, d_output, index, dropout=0.1, rng=None, site="tdnn"):
        super().__init__()
        self.concat = ConcatLayer(index)
        self.proj = nn.Linear(d_input * len(index), d_output, bias=True)
        init.xavier_normal_(self.proj.weight)
        self.p = float(dropout)
        self._rng = rng or _rng.GLOBAL
        self._site = self._rng.site(site)

    def forward(self, x):
        drop = self._rng.make(self.p, self._site, x.device, self.training)
        if x.dtype == torch.bfloat16:          # bf16 path: n_ctx shifted TMA boxes accumulate into one TMEM tile
            return ops.linear_tc(x, self.proj.weight, self.proj.bias, splice=self.concat.index, relu=True, drop=drop)
        return ops.linear(x, self.proj.weight, self.proj.bias, splice=self.concat.index, relu=True, drop=drop)


class LDALayer(nn.Module):
    """Frozen affine transform from a Kaldi lda.mat: y = x @ W + b, W = lda[:, :-1]^T, b = lda[:, -1]
    (L/pytorch/TDNN.py:48-55).  No gradient flows into it or through it (its input is the feature tensor)."""

    def __init__(self, LDA_mat):
        super().__init__()
        mat = torch.as_tensor(LDA_mat, dtype=torch.float32)
        self.weight = nn.Parameter(mat[:, :-1].t().contiguous(), requires_grad=False)     # [in, out]
        self.bias = nn.Parameter(mat[:, -1].contiguous(), requires_grad=False)

    def forward(self, x):
        if x.dtype == torch.bfloat16:
            # frozen matrix: its bf16 [out, in] operand copy is made once and lives with the module
            key = (self.weight._version, self.weight.data_ptr())
            if getattr(self, "_wt_key", None) != key:
                self._wt, self._wt_key = ops.transpose_to_bf16(self.weight), key
            return ops.affine_tc(x, self.weight, self.bias, wt=self._wt)
        return ops.affine_kn(x, self.weight, self.bias)
