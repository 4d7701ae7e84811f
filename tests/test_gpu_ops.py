"""Per-kernel parity (B200, `-m gpu`): every op of libpka_b200.so, forward and backward, against the CPU oracle
(oracle/acoustic_model.py) / plain torch fp32 on the same seeded inputs, through the C ABI.

Stated tolerances (fp32 path): forward values rtol 1e-4 / atol 1e-5 relative to the tensor scale;
gradients 1e-3 relative to the per-tensor max-abs.  Index / integer outputs are exact.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import acoustic_model as am          # noqa: E402
from oracle import cmvn as ocmvn                 # noqa: E402
from oracle import train_step as otrain          # noqa: E402


def P():
    import pytorch_kaldi_asr_b200 as pk
    return pk


def ops():
    from pytorch_kaldi_asr_b200 import ops as o
    return o


DEV = "cuda"


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-6))


def assert_close(a, b, tol, what=""):
    e = rel_err(a, b)
    assert e <= tol, "%s: relative error %.3e > %.1e" % (what, e, tol)


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


# ------------------------------------------------------------------------------------------------ library
def test_library_loads_on_sm100():
    from pytorch_kaldi_asr_b200 import _lib
    assert _lib.lib().pka_check_device() == 0, _lib.lib().pka_last_error()


# ------------------------------------------------------------------------------------------------ GEMM family
@pytest.mark.parametrize("B,T,kin,N,ctx,relu", [
    (3, 17, 40, 24, None, False),          # plain linear, odd sizes
    (2, 50, 32, 256, [-1, 0, 1], True),    # TDNN, small tiles
    (32, 430, 256, 256, [-3, 0, 3], True), # TDNN at TIMIT size (big-tile config)
    (4, 23, 200, 256, None, False),        # src_projection shape
    (4, 9, 128, 53, None, False),          # vocabulary projection (N not a multiple of 4)
])
def test_linear_fwd_bwd(B, T, kin, N, ctx, relu):
    o = ops()
    n_ctx = len(ctx) if ctx else 1
    x = rnd(B, T, kin, seed=1)
    w = rnd(N, n_ctx * kin, seed=2, scale=1.0 / math.sqrt(n_ctx * kin))
    b = rnd(N, seed=3, scale=0.1)
    gy = rnd(B, T, N, seed=4)
    xg, wg, bg = [t.clone().to(DEV).requires_grad_(True) for t in (x, w, b)]
    out = o.linear(xg, wg, bg, splice=ctx, relu=relu)
    out.backward(gy.to(DEV))
    xr, wr, br = [t.clone().requires_grad_(True) for t in (x, w, b)]
    ref = F.linear(am.splice(xr, ctx) if ctx else xr, wr, br)
    if relu:
        # the ReLU gate of a pre-activation that is zero to rounding (|z| ~ 1e-7) may legitimately differ between two
        # fp32 summation orders; take the gate from the kernel output, after checking it only differs at such points
        gate = out.detach().cpu() > 0
        flips = gate != (ref.detach() > 0)
        assert float(ref.detach()[flips].abs().max()) < 1e-5 if flips.any() else True
        ref = ref * gate
    ref.backward(gy)
    assert_close(out, ref, 1e-4, "linear fwd")
    assert_close(xg.grad, xr.grad, 1e-3, "linear dx")
    assert_close(wg.grad, wr.grad, 1e-3, "linear dW")
    assert_close(bg.grad, br.grad, 1e-3, "linear db")


def test_linear_dropout_residual_matches_injected_mask():
    o = ops()
    B, T, kin, N, p = 3, 11, 64, 128, 0.35
    x, w, b, r = rnd(B, T, kin, seed=1), rnd(N, kin, seed=2, scale=0.1), rnd(N, seed=3), rnd(B, T, N, seed=5)
    step = torch.tensor([7], dtype=torch.int64, device=DEV)
    drop = o.Drop(p, 42, 1234, step)
    keep = o.dropout_keep_mask(B * T * N, drop, DEV).cpu().view(B, T, N).float()
    assert 0.55 < keep.mean().item() < 0.75
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    ref = F.linear(xr, wr, b) * keep / (1 - p) + r
    gy = rnd(B, T, N, seed=6)
    ref.backward(gy)
    xg, wg, rg = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True), r.to(DEV).requires_grad_(True)
    out = o.linear(xg, wg, b.to(DEV), drop=drop, residual=rg)
    out.backward(gy.to(DEV))
    assert_close(out, ref, 1e-4, "fwd")
    assert_close(xg.grad, xr.grad, 1e-3, "dx")
    assert_close(wg.grad, wr.grad, 1e-3, "dW")
    assert_close(rg.grad, gy, 1e-6, "dres")


def test_head_proj_fwd_bwd():
    o = ops()
    B, L, D, H, dk = 3, 19, 128, 2, 64
    x = rnd(B, L, D, seed=1)
    ws = [rnd(H, D, dk, seed=10 + i, scale=0.1) for i in range(3)]
    gy = rnd(B, L, 3 * H * dk, seed=5)
    xr = x.clone().requires_grad_(True)
    wr = [w.clone().requires_grad_(True) for w in ws]
    ref = torch.cat([torch.einsum("bld,hdk->blhk", xr, w).reshape(B, L, H * dk) for w in wr], dim=-1)
    ref.backward(gy)
    xg = x.to(DEV).requires_grad_(True)
    wg = [w.to(DEV).requires_grad_(True) for w in ws]
    out = o.head_proj(xg, *wg)
    out.backward(gy.to(DEV))
    assert_close(out, ref, 1e-4, "head_proj fwd")
    assert_close(xg.grad, xr.grad, 1e-3, "head_proj dx")
    for a, b in zip(wg, wr):
        assert_close(a.grad, b.grad, 1e-3, "head_proj dw")


def test_lda_affine():
    o = ops()
    x, w, b = rnd(2, 13, 200, seed=1), rnd(200, 200, seed=2, scale=0.1), rnd(200, seed=3)
    out = o.affine_kn(x.to(DEV), w.to(DEV), b.to(DEV))
    assert_close(out, x.matmul(w) + b, 1e-4, "lda")


# ------------------------------------------------------------------------------------------------ attention
def dense_attention(q, k, v, key_mask, band, scale, keep=None, p=0.0):
    """q [B,H,Lq,D] etc; returns out [B,Lq,H*D], probs."""
    B, H, Lq, D = q.shape
    masked = am.attention_mask(Lq, key_mask, band)[:, None]
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    dead = masked.all(-1, keepdim=True)
    s = torch.where(dead, torch.zeros_like(s), s.masked_fill(masked, float("-inf")))
    pr = torch.softmax(s, -1).masked_fill(masked, 0.0)
    a = pr if keep is None else pr * keep / (1 - p)
    out = torch.matmul(a, v).permute(0, 2, 1, 3).reshape(B, Lq, H * D)
    return out, pr


@pytest.mark.parametrize("B,H,Lq,Lk,D,band,selfattn,p", [
    (3, 2, 21, 21, 64, (-10, 0), True, 0.0),     # decoder self-attention band
    (2, 2, 55, 55, 64, (-3, 1), True, 0.0),
    (3, 2, 21, 70, 64, None, False, 0.0),        # cross-attention, key padding only
    (2, 3, 9, 40, 16, None, False, 0.0),
    (2, 2, 33, 33, 32, (-4, 0), True, 0.35),     # dropout on the probabilities, injected mask
    (2, 8, 150, 150, 64, (-100, 0), True, 0.0),  # config-5 style band
])
def test_attention_fwd_bwd(B, H, Lq, Lk, D, band, selfattn, p):
    o = ops()
    HD = H * D
    scale = 1.0 / math.sqrt(128.0)
    lens = torch.tensor([Lk, max(1, Lk // 2), max(1, Lk - 3)][:B])
    key_mask = (torch.arange(Lk)[None, :] < lens[:, None]).to(torch.uint8)
    if selfattn:
        qkv = rnd(B, Lq, 3 * HD, seed=1)
        q, k, v = [t.reshape(B, Lq, H, D).permute(0, 2, 1, 3) for t in qkv.split(HD, dim=-1)]
        qbuf, kvbuf = qkv, None
    else:
        qbuf, kvbuf = rnd(B, Lq, HD, seed=1), rnd(B, Lk, 2 * HD, seed=2)
        q = qbuf.reshape(B, Lq, H, D).permute(0, 2, 1, 3)
        k, v = [t.reshape(B, Lk, H, D).permute(0, 2, 1, 3) for t in kvbuf.split(HD, dim=-1)]
    keep, drop = None, None
    if p > 0:
        drop = o.Drop(p, 5, 99, torch.tensor([3], dtype=torch.int64, device=DEV))
        keep = o.attn_keep_mask(B, H, Lq, Lk, drop, DEV).cpu().float()
    qr = qbuf.clone().requires_grad_(True)
    kvr = kvbuf.clone().requires_grad_(True) if kvbuf is not None else None
    if selfattn:
        q_, k_, v_ = [t.reshape(B, Lq, H, D).permute(0, 2, 1, 3) for t in qr.split(HD, dim=-1)]
    else:
        q_ = qr.reshape(B, Lq, H, D).permute(0, 2, 1, 3)
        k_, v_ = [t.reshape(B, Lk, H, D).permute(0, 2, 1, 3) for t in kvr.split(HD, dim=-1)]
    ref, ref_probs = dense_attention(q_, k_, v_, key_mask, band, scale, keep, p)
    gy = rnd(B, Lq, HD, seed=7)
    ref.backward(gy)
    qg = qbuf.to(DEV).requires_grad_(True)
    kvg = kvbuf.to(DEV).requires_grad_(True) if kvbuf is not None else None
    out, probs = o.attention(qg, kvg, key_mask.to(DEV), H, D, band, scale, drop, want_probs=True)
    out.backward(gy.to(DEV))
    assert_close(out, ref, 1e-4, "attn fwd")
    assert_close(probs, ref_probs, 1e-4, "attn probs")
    assert_close(qg.grad, qr.grad, 1e-3, "attn dq(kv)")
    if kvbuf is not None:
        assert_close(kvg.grad, kvr.grad, 1e-3, "attn dkv")


@pytest.mark.parametrize("B,H,Lq,Lk", [(3, 2, 10, 499), (2, 2, 3, 64), (3, 1, 16, 131), (2, 3, 7, 700), (2, 2, 12, 1024)])
def test_attention_few_queries_many_keys(B, H, Lq, Lk):
    """The beam-search cross-attention shape (<= 16 queries, head width 64, no band, no probabilities requested): the
    one-pass kernel attn_fwd_smallq_kernel against dense attention.  Masks with holes, an utterance whose keys are all
    masked (output 0, no NaN)."""
    o = ops()
    D, HD = 64, H * 64
    scale = 1.0 / math.sqrt(128.0)
    g = torch.Generator().manual_seed(Lk)
    key_mask = (torch.rand(B, Lk, generator=g) > 0.3).to(torch.uint8)
    key_mask[0, Lk // 3:] = 0
    key_mask[0, :2] = 1
    if B > 2:
        key_mask[2] = 0                                        # dead utterance
    qbuf, kvbuf = rnd(B, Lq, HD, seed=1), rnd(B, Lk, 2 * HD, seed=2)
    q_ = qbuf.reshape(B, Lq, H, D).permute(0, 2, 1, 3)
    k_, v_ = [t.reshape(B, Lk, H, D).permute(0, 2, 1, 3) for t in kvbuf.split(HD, dim=-1)]
    ref, _ = dense_attention(q_, k_, v_, key_mask, None, scale, None, 0.0)
    out, _ = o.attention(qbuf.to(DEV), kvbuf.to(DEV), key_mask.to(DEV), H, D, None, scale, None)
    assert torch.isfinite(out).all()
    assert_close(out, ref, 1e-4, "few-query attention")
    if B > 2:
        assert float(out[2].abs().max()) == 0.0


def test_attention_fully_masked_rows_are_zero_not_nan():
    o = ops()
    B, H, L, D = 2, 2, 12, 64
    qkv = rnd(B, L, 3 * H * D, seed=1).to(DEV).requires_grad_(True)
    key_mask = torch.zeros(B, L, dtype=torch.uint8)
    key_mask[0, :5] = 1                    # batch 1 has no real key at all; batch 0 rows >= 5+band see none either
    out, _ = o.attention(qkv, None, key_mask.to(DEV), H, D, (-2, 0), 0.1, None)
    out.sum().backward()
    assert torch.isfinite(out).all() and torch.isfinite(qkv.grad).all()
    assert float(out[1].abs().max()) == 0.0 and float(qkv.grad[1].abs().max()) == 0.0
    assert float(out[0, 8:].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------ add + LayerNorm
# (lane-group geometries of csrc/layernorm.cu: 8 / 16 / 32 lanes per row, 1-4 vectors per lane, widths that do not fill the
# group, row counts that leave a warp's last rows empty; the last two walk several rows per warp behind the prefetch with
# 296 and 888 partial rows)
@pytest.mark.parametrize("rows,D,p", [(40, 32, 0.0), (1760, 128, 0.0), (300, 256, 0.35), (76, 512, 0.35), (44, 40, 0.35),
                                      (52, 200, 0.0), (36, 64, 0.35), (10004, 128, 0.35), (30000, 64, 0.0)])
def test_add_layernorm_fwd_bwd(rows, D, p):
    o = ops()
    x, r = rnd(4, rows // 4, D, seed=1), rnd(4, rows // 4, D, seed=2)
    a, b = rnd(D, seed=3) * 0.5 + 1.0, rnd(D, seed=4) * 0.1
    gy = rnd(4, rows // 4, D, seed=5)
    keep, drop = None, None
    if p > 0:
        drop = o.Drop(p, 11, 5, torch.tensor([2], dtype=torch.int64, device=DEV))
        keep = o.dropout_keep_mask(x.numel(), drop, DEV).cpu().view(x.shape).float()
    xr, rr, ar, br = [t.clone().requires_grad_(True) for t in (x, r, a, b)]
    z = (xr if keep is None else xr * keep / (1 - p)) + rr
    ref = am.layer_norm_ref(z, ar, br)
    ref.backward(gy)
    xg, rg, ag, bg = [t.clone().to(DEV).requires_grad_(True) for t in (x, r, a, b)]
    out = o.add_layer_norm(xg, rg, ag, bg, 1e-3, drop)
    out.backward(gy.to(DEV))
    assert_close(out, ref, 1e-4, "ln fwd")
    assert_close(xg.grad, xr.grad, 1e-3, "ln dx")
    assert_close(rg.grad, rr.grad, 1e-3, "ln dres")
    assert_close(ag.grad, ar.grad, 1e-3, "ln da")
    assert_close(bg.grad, br.grad, 1e-3, "ln db")


@pytest.mark.parametrize("rows,D,p", [(2016, 128, 0.35), (300, 256, 0.0), (52, 512, 0.1), (44, 64, 0.0), (36, 200, 0.35),
                                      (36, 204, 0.0), (12792, 512, 0.1), (40000, 256, 0.0)])
def test_add_layernorm_bf16_fwd_bwd(rows, D, p):
    """The bf16 instantiations (16-byte vectors of 8, 8-byte vectors of 4 for D % 8 != 0) against fp32 math on the same
    bf16-rounded inputs: outputs and data gradients within bf16 rounding (1e-2 of the tensor scale), the fp32 gain /
    offset gradients 2e-3."""
    o = ops()
    bf = lambda t: t.bfloat16().float()
    x, r = bf(rnd(4, rows // 4, D, seed=1)), bf(rnd(4, rows // 4, D, seed=2))
    a, b = rnd(D, seed=3) * 0.5 + 1.0, rnd(D, seed=4) * 0.1
    gy = bf(rnd(4, rows // 4, D, seed=5))
    keep, drop = None, None
    if p > 0:
        drop = o.Drop(p, 11, 5, torch.tensor([2], dtype=torch.int64, device=DEV))
        keep = o.dropout_keep_mask(x.numel(), drop, DEV).cpu().view(x.shape).float()
    xr, rr, ar, br = [t.clone().requires_grad_(True) for t in (x, r, a, b)]
    z = (xr if keep is None else xr * keep / (1 - p)) + rr
    ref = am.layer_norm_ref(z, ar, br)
    ref.backward(gy)
    xg, rg = [t.clone().to(DEV).bfloat16().requires_grad_(True) for t in (x, r)]
    ag, bg = [t.clone().to(DEV).requires_grad_(True) for t in (a, b)]
    out = o.add_layer_norm(xg, rg, ag, bg, 1e-3, drop)
    assert out.dtype == torch.bfloat16
    out.backward(gy.to(DEV).bfloat16())
    assert_close(out.float(), ref, 1e-2, "ln bf16 fwd")
    assert_close(xg.grad.float(), xr.grad, 1e-2, "ln bf16 dx")
    assert_close(rg.grad.float(), rr.grad, 1e-2, "ln bf16 dres")
    assert_close(ag.grad, ar.grad, 2e-3, "ln bf16 da")
    assert_close(bg.grad, br.grad, 2e-3, "ln bf16 db")


# ------------------------------------------------------------------------------------------------ loss
@pytest.mark.parametrize("smoothing", [False, True])
@pytest.mark.parametrize("N,V", [(37, 11), (1760, 53)])
def test_cross_entropy(N, V, smoothing):
    o = ops()
    lg = rnd(N, V, seed=1) * 3
    goal = torch.randint(0, V, (N,), generator=torch.Generator().manual_seed(2))
    goal[::5] = 0
    lr = lg.clone().requires_grad_(True)
    ref = am.cross_entropy_sum(lr, goal, smoothing)
    (ref * 0.7).backward()
    lgd = lg.to(DEV).requires_grad_(True)
    loss, stats = o.cross_entropy_sum(lgd, goal.to(DEV), smoothing)
    (loss * 0.7).backward()
    assert_close(loss, ref, 2e-5, "ce loss")
    assert_close(lgd.grad, lr.grad, 1e-4, "ce grad")
    keep = goal.ne(0)
    assert int(stats[1].item()) == int(keep.sum())
    assert int(stats[0].item()) == int((lg.argmax(1).eq(goal) & keep).sum())


# ------------------------------------------------------------------------------------------------ embedding / positions
def test_embed_pos_fwd_bwd():
    o = ops()
    B, L, D, V = 4, 13, 128, 53
    emb = rnd(V, D, seed=1)
    emb[0] = 0
    pos = am.sinusoid_table(100, D)
    tok = torch.randint(0, V, (B, L), generator=torch.Generator().manual_seed(3))
    er = emb.clone().requires_grad_(True)
    ref = er[tok] + pos[:L][None]
    gy = rnd(B, L, D, seed=4)
    ref.backward(gy)
    eg = emb.to(DEV).requires_grad_(True)
    out = o.embed_pos(tok.to(DEV), eg, pos.to(DEV))
    out.backward(gy.to(DEV))
    assert_close(out, ref, 1e-6, "embed fwd")
    want = er.grad.clone()
    want[0] = 0                                   # padding_idx row receives no gradient (nn.Embedding(padding_idx=0))
    assert_close(eg.grad, want, 1e-5, "embed bwd")


def test_add_pos_dropout():
    o = ops()
    B, T, D, p = 3, 20, 256, 0.35
    x, pos = rnd(B, T, D, seed=1), am.sinusoid_table(50, D)
    drop = o.Drop(p, 8, 77, torch.tensor([5], dtype=torch.int64, device=DEV))
    keep = o.dropout_keep_mask(B * T * D, drop, DEV).cpu().view(B, T, D).float()
    xg = x.to(DEV).requires_grad_(True)
    out = o.add_pos_dropout(xg, pos.to(DEV), drop)
    gy = rnd(B, T, D, seed=2)
    out.backward(gy.to(DEV))
    assert_close(out, (x + pos[:T][None]) * keep / (1 - p), 1e-6, "add_pos_dropout fwd")
    assert_close(xg.grad, gy * keep / (1 - p), 1e-6, "add_pos_dropout bwd")


def test_dropout_mask_depends_on_step_and_site_and_rate():
    o = ops()
    n = 1 << 16
    st = torch.tensor([1], dtype=torch.int64, device=DEV)
    a = o.dropout_keep_mask(n, o.Drop(0.35, 3, 9, st), DEV)
    b = o.dropout_keep_mask(n, o.Drop(0.35, 4, 9, st), DEV)
    a2 = o.dropout_keep_mask(n, o.Drop(0.35, 3, 9, st), DEV)
    st.add_(1)
    c = o.dropout_keep_mask(n, o.Drop(0.35, 3, 9, st), DEV)
    assert torch.equal(a, a2) and not torch.equal(a, b) and not torch.equal(a, c)
    assert abs(a.float().mean().item() - 0.65) < 0.01


# ------------------------------------------------------------------------------------------------ front-end
@pytest.mark.parametrize("fold,cmvn", [(1, 0), (2, 0), (1, 1), (3, 2)])
@pytest.mark.parametrize("T,Fd,ctx", [(37, 40, [-2, -1, 0, 1, 2]), (150, 40, [-2, -1, 0, 1, 2]), (203, 24, [-3, 0, 4]),
                                      (131, 6, [0]), (90, 40, [1, 2])])
def test_frontend_fold_splice_cmvn(fold, cmvn, T, Fd, ctx):
    """T > 64 spans several shared-memory tiles of the tile kernel (halo frames across tile edges, a last partial tile);
    F = 6 takes the direct kernel (no 16-byte vectors); one-sided and asymmetric contexts."""
    o = ops()
    B = 4
    lens = np.array([T, T // 2 + 1, T - 6, 5])
    x = rnd(B, T, Fd, seed=1).numpy() * 2 + 0.5
    for b, n in enumerate(lens):
        x[b, n:] = 0
    ref = torch.from_numpy(ocmvn.apply_cmvn(x, lens, norm_vars=(cmvn == 2)) if cmvn else x)
    mask = torch.from_numpy((np.arange(T)[None] < lens[:, None]).astype(np.uint8))
    ref, _ = am.fold_frames(ref, mask, fold)
    ref = am.splice(ref, ctx)
    out = o.frontend(torch.from_numpy(x).to(DEV), torch.from_numpy(lens).to(DEV), fold, ctx, cmvn)
    assert out.shape == ref.shape
    assert_close(out, ref, 1e-5 if cmvn else 0.0, "frontend")
    out16 = o.frontend(torch.from_numpy(x).to(DEV), torch.from_numpy(lens).to(DEV), fold, ctx, cmvn, out_dtype=torch.bfloat16)
    assert out16.dtype == torch.bfloat16 and out16.shape == ref.shape
    assert torch.equal(out16, out.bfloat16()), "frontend bf16: not the fp32 result rounded once"


@pytest.mark.parametrize("cmvn", [1, 2])
def test_frontend_cmvn_vs_hand_computed_vectors(cmvn):
    """The CMVN of the fused front-end kernel against the hand-computed known-answer vectors
    (tests/golden/cmvn_hand_computed.json, derivation inside): all-equal feature (variance floor), N = 1, N = 0."""
    import json
    import os
    o = ops()
    with open(os.path.join(os.path.dirname(__file__), "golden", "cmvn_hand_computed.json")) as fh:
        g = json.load(fh)
    x = torch.tensor(g["feats"], dtype=torch.float32)
    lens = torch.tensor(g["lengths"], dtype=torch.int32)
    want = torch.tensor(g["mean_var" if cmvn == 2 else "mean_only"], dtype=torch.float32)
    out = o.frontend(x.to(DEV), lens.to(DEV), 1, [0], cmvn)
    assert out.shape == want.shape
    assert (out.cpu() - want).abs().max().item() <= 1e-6


def test_concat_layer_known_vector():
    pk = P()
    from pytorch_kaldi_asr_b200.TDNN import ConcatLayer
    x = torch.arange(1, 5, dtype=torch.float32).view(1, 4, 1).to(DEV)
    got = ConcatLayer([-2, -1, 0, 1, 2])(x).cpu()[0].tolist()
    assert got == [[0, 0, 1, 2, 3], [0, 1, 2, 3, 4], [1, 2, 3, 4, 0], [2, 3, 4, 0, 0]]


# ------------------------------------------------------------------------------------------------ optimiser
def test_fused_adam_matches_oracle_adam_and_schedule():
    pk = P()
    torch.manual_seed(0)
    ps = [torch.randn(33, 17), torch.randn(5), torch.randn(2, 8, 16)]
    gs = [[torch.randn_like(p) * (0.1 + i) for p in ps] for i in range(5)]
    ref = {str(i): p.clone() for i, p in enumerate(ps)}
    osch = otrain.AdamSchedule(ref, start_lr=2e-3, soft_coefficient=10)
    params = [torch.nn.Parameter(p.clone().to(DEV)) for p in ps]
    opt = pk.ScheduledOptim(pk.FusedAdam(params, betas=(0.9, 0.999), eps=1e-8), 2e-3, 10)
    for step in range(5):
        osch.step({str(i): g for i, g in enumerate(gs[step])})
        osch.update_learning_rate()
        if step % 2 == 0:                      # classic semantics: the arena is cleared, gradients are copied into the views
            opt.optimizer.zero_grad(set_to_none=False)
            for p, g in zip(params, gs[step]):
                p.grad.copy_(g.to(DEV))
        else:                                  # default: .grad dropped; a foreign gradient tensor is adopted into the arena
            opt.zero_grad()
            assert all(p.grad is None for p in params)
            for p, g in zip(params, gs[step]):
                p.grad = g.to(DEV)
        opt.step()
        opt.update_learning_rate()
    for i, p in enumerate(params):
        assert_close(p.data, ref[str(i)], 2e-6, "adam param %d" % i)
    assert opt.n_current_steps == 5
    assert int(opt.optimizer.dev_state[0]) == 5 and int(opt.optimizer.dev_state[1]) == 5
    assert abs(float(opt.optimizer.dev_lr[0]) - osch.lr) < 1e-9


@pytest.mark.parametrize("p,site,seed,step,n", [(0.35, 3, 1234, 7, 100003), (0.1, 24, (0xdeadbeef << 32) | 17, (5 << 32) | 9, 4099),
                                                (0.999, 1, 1, 1, 257), (0.5, 0, 0, 0, 64)])
def test_dropout_bits_equal_the_numpy_philox_restatement(p, site, seed, step, n):
    """The keep bits every kernel derives (pka_dropout_mask materialises them) are bit-identical to oracle/philox.py,
    which reproduces the published Philox4x32-10 known-answer vectors (CPU test)."""
    from oracle import philox
    o = ops()
    st = torch.tensor([step], dtype=torch.int64, device=DEV)
    got = o.dropout_keep_mask(n, o.Drop(p, site, seed, st), DEV).cpu().numpy()
    assert np.array_equal(got, philox.keep_mask(n, p, site, seed, step))
