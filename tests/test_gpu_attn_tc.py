"""Tensor-core attention (tcgen05 / TMEM / TMA, csrc/attn_tc.cu) parity on the B200 (`-m gpu`), through the C ABI.

Checker: the oracle's restatement of ScaledDotProductAttention + masks (T/Modules.py:75-97, T/Models.py:27-49) in fp32
torch on the bf16-rounded q/k/v the kernel actually multiplies.  Stated tolerance: outputs within 1e-2 of the output
scale (the probabilities are rounded to bf16 before the P.V product: relative 2^-9 per term), lse within 2e-3 absolute;
dead rows exactly 0 / -inf.  With dropout the kernel's Philox keep bits are materialised and injected in the checker."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def ref_attention(q, k, v, key_mask, band, scale, keep=None, drop_scale=1.0):
    """q [B,H,Lq,D] k,v [B,H,Lk,D] fp32; key_mask [B,Lk] (1 = real) -> out [B,H,Lq,D], lse [B,H,Lq]."""
    B, H, Lq, D = q.shape
    Lk = k.shape[2]
    s = torch.einsum("bhid,bhjd->bhij", q, k) * scale
    masked = key_mask.eq(0)[:, None, None, :].expand(B, H, Lq, Lk).clone()
    if band is not None:
        i = torch.arange(Lq)[:, None]
        j = torch.arange(Lk)[None, :]
        masked |= ((j < i + band[0]) | (j > i + band[1]))[None, None]
    s = s.masked_fill(masked, float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    p = torch.softmax(s, dim=-1).masked_fill(masked, 0.0)
    p = torch.nan_to_num(p, nan=0.0)
    if keep is not None:
        p = p * keep * drop_scale
    return torch.einsum("bhij,bhjd->bhid", p, v), lse


CASES = [
    # B, H, Lq, Lk, band, pad_keys, self
    (2, 2, 63, 63, (-10, 0), 5, True),           # TIMIT decoder self-attention (band -10..0)
    (3, 2, 63, 499, None, 120, False),           # TIMIT cross-attention, ragged key padding, 4 key tiles
    (2, 8, 300, 300, (-100, 0), 17, True),       # cfg5-style banded encoder self-attention, 3 query tiles
    (1, 4, 257, 257, None, 0, True),             # full attention, tile tails (257 = 2*128 + 1)
    (2, 2, 130, 140, (-3, 2), 139, False),       # almost everything padded: utterance 1 has a single real key
    (1, 1, 128, 128, (5, 9), 0, True),           # band entirely in the future: trailing rows have no allowed key
]


@pytest.mark.parametrize("B,H,Lq,Lk,band,pad,self_attn", CASES)
@pytest.mark.parametrize("out_fp32", [False, True])
def test_attn_tc_forward_matches_oracle(B, H, Lq, Lk, band, pad, self_attn, out_fp32):
    from pytorch_kaldi_asr_b200 import ops
    D, HD = 64, H * 64
    scale = 1.0 / math.sqrt(HD * 2)
    if self_attn:
        qkv = rnd(B, Lq, 3 * HD, seed=1).bfloat16()
        qbuf, kvbuf = qkv.to(DEV), None
        q, k, v = qkv[..., :HD], qkv[..., HD:2 * HD], qkv[..., 2 * HD:]
    else:
        qb = rnd(B, Lq, HD, seed=1).bfloat16()
        kv = rnd(B, Lk, 2 * HD, seed=2).bfloat16()
        qbuf, kvbuf = qb.to(DEV), kv.to(DEV)
        q, k, v = qb, kv[..., :HD], kv[..., HD:]
    key_mask = torch.ones(B, Lk, dtype=torch.uint8)
    for b in range(B):
        n_pad = (pad * (b + 1)) // B if pad else 0
        if n_pad:
            key_mask[b, Lk - n_pad:] = 0
    out, lse = ops.attention_tc(qbuf, kvbuf, key_mask.to(DEV), H, D, band, scale, None, out_fp32)
    assert out.dtype == (torch.float32 if out_fp32 else torch.bfloat16)
    split = lambda t, L_: t.float().view(B, L_, H, D).permute(0, 2, 1, 3)
    ref, lse_ref = ref_attention(split(q, Lq), split(k, Lk), split(v, Lk), key_mask, band, scale)
    ref = ref.permute(0, 2, 1, 3).reshape(B, Lq, HD)
    got = out.float().cpu()
    tol = 1e-2 * float(ref.abs().max())
    assert float((got - ref).abs().max()) <= tol, (float((got - ref).abs().max()), tol)
    dead = torch.isinf(lse_ref)
    lse_c = lse.cpu()
    assert torch.equal(torch.isinf(lse_c), dead)
    assert float((lse_c[~dead] - lse_ref[~dead]).abs().max()) <= 2e-3
    dead_rows = dead.permute(0, 2, 1)[..., None].expand(B, Lq, H, D).reshape(B, Lq, HD)
    assert float(got[dead_rows].abs().max()) == 0.0 if dead_rows.any() else True


def test_attn_tc_dropout_uses_the_shared_philox_bits():
    from pytorch_kaldi_asr_b200 import ops
    B, H, Lq, Lk, D = 2, 2, 63, 200, 64
    HD = H * D
    scale = 1.0 / math.sqrt(128.0)
    qb = rnd(B, Lq, HD, seed=3).bfloat16()
    kv = rnd(B, Lk, 2 * HD, seed=4).bfloat16()
    key_mask = torch.ones(B, Lk, dtype=torch.uint8)
    key_mask[1, 150:] = 0
    step = torch.full((1,), 7, dtype=torch.int64, device=DEV)
    drop = ops.Drop(0.35, 11, 1234, step)
    out, _ = ops.attention_tc(qb.to(DEV), kv.to(DEV), key_mask.to(DEV), H, D, None, scale, drop, True)
    keep = ops.attn_keep_mask(B, H, Lq, Lk, drop, DEV).float().cpu()
    split = lambda t, L_: t.float().view(B, L_, H, D).permute(0, 2, 1, 3)
    ref, _ = ref_attention(split(qb, Lq), split(kv[..., :HD], Lk), split(kv[..., HD:], Lk), key_mask, None, scale,
                           keep=keep, drop_scale=1.0 / 0.65)
    ref = ref.permute(0, 2, 1, 3).reshape(B, Lq, HD)
    got = out.cpu()
    assert float((got - ref).abs().max()) <= 1e-2 * float(ref.abs().max())


BWD_CASES = [
    # B, H, Lq, Lk, band, pad, self, dropout
    (2, 2, 63, 300, None, 50, False, 0.0),        # TIMIT-like cross-attention
    (2, 2, 63, 63, (-10, 0), 7, True, 0.0),       # decoder self-attention band
    (2, 4, 300, 300, (-100, 0), 33, True, 0.0),   # banded encoder self-attention, several tiles both ways
    (1, 2, 200, 333, None, 0, False, 0.35),       # dropout: backward regenerates the forward's Philox bits
    (1, 2, 150, 150, (-20, 5), 10, True, 0.35),
    (1, 1, 128, 128, (5, 9), 0, True, 0.0),       # rows without any allowed key: zero gradient
]


@pytest.mark.parametrize("B,H,Lq,Lk,band,pad,self_attn,pdrop", BWD_CASES)
@pytest.mark.parametrize("tc_bwd", [True, False])
def test_attn_tc_backward_matches_autograd_of_the_oracle(B, H, Lq, Lk, band, pad, self_attn, pdrop, tc_bwd):
    """Gradients of the bf16 path (tensor-core forward + tensor-core / SIMT backward kernels) against autograd through
    the fp32 checker on the same bf16-rounded inputs.  Stated tolerance: 2e-2 of each gradient tensor's scale."""
    from pytorch_kaldi_asr_b200 import ops
    D, HD = 64, H * 64
    scale = 1.0 / math.sqrt(128.0)
    key_mask = torch.ones(B, Lk, dtype=torch.uint8)
    for b in range(B):
        n_pad = (pad * (b + 1)) // B if pad else 0
        if n_pad:
            key_mask[b, Lk - n_pad:] = 0
    gy = rnd(B, Lq, HD, seed=7).bfloat16()
    drop, keep = None, None
    if pdrop > 0:
        drop = ops.Drop(pdrop, 5, 4321, torch.full((1,), 3, dtype=torch.int64, device=DEV))
        keep = ops.attn_keep_mask(B, H, Lq, Lk, drop, DEV).float().cpu()
    split = lambda t, L_: t.view(B, L_, H, D).permute(0, 2, 1, 3)
    old = ops.ATTN_BWD_TC
    ops.ATTN_BWD_TC = tc_bwd
    try:
        if self_attn:
            qkv = rnd(B, Lq, 3 * HD, seed=5).bfloat16()
            xg = qkv.to(DEV).requires_grad_(True)
            out, _ = ops.attention_tc(xg, None, key_mask.to(DEV), H, D, band, scale, drop, False)
            out.backward(gy.to(DEV))
            xr = qkv.float().requires_grad_(True)
            ref, _ = ref_attention(split(xr[..., :HD], Lq), split(xr[..., HD:2 * HD], Lk), split(xr[..., 2 * HD:], Lk), key_mask,
                                   band, scale, keep=keep, drop_scale=1.0 / (1.0 - pdrop))
            pairs = [(xg, xr)]
        else:
            qb = rnd(B, Lq, HD, seed=5).bfloat16()
            kv = rnd(B, Lk, 2 * HD, seed=6).bfloat16()
            qg, kvg = qb.to(DEV).requires_grad_(True), kv.to(DEV).requires_grad_(True)
            out, _ = ops.attention_tc(qg, kvg, key_mask.to(DEV), H, D, band, scale, drop, False)
            out.backward(gy.to(DEV))
            qr, kvr = qb.float().requires_grad_(True), kv.float().requires_grad_(True)
            ref, _ = ref_attention(split(qr, Lq), split(kvr[..., :HD], Lk), split(kvr[..., HD:], Lk), key_mask, band, scale,
                                   keep=keep, drop_scale=1.0 / (1.0 - pdrop))
            pairs = [(qg, qr), (kvg, kvr)]
    finally:
        ops.ATTN_BWD_TC = old
    ref.permute(0, 2, 1, 3).reshape(B, Lq, HD).backward(gy.float())
    for got, want in pairs:
        g, w = got.grad.float().cpu(), want.grad
        assert torch.isfinite(g).all()
        assert float((g - w).abs().max()) <= 2e-2 * float(w.abs().max()), (float((g - w).abs().max()), float(w.abs().max()))
