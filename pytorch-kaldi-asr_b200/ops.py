"""torch.autograd.Function wrappers over the C ABI of libpka_b200.so.

Every function here launches hand-written sm_100a kernels on torch's current stream; torch only provides memory,
streams and the autograd tape.  Nothing falls back to ATen: non-CUDA tensors raise.

Precision modes (module-level switch, see `set_compute_mode`):
  "fp32": every GEMM runs the fp32 SIMT kernel with a fixed summation order (parity / exact beam search).
  "bf16": GEMMs whose shapes fit the tcgen05 tiles run on the tensor cores with bf16 operands (fp32 master
          weights, fp32 accumulation in TMEM); everything else stays fp32.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading
from typing import Optional, Sequence

import torch

from . import _lib as L

_MODE = {"gemm": "fp32"}


def set_compute_mode(mode: str):
    assert mode in ("fp32", "bf16")
    _MODE["gemm"] = mode


def compute_mode() -> str:
    return _MODE["gemm"]


class Drop:
    """One dropout call site of the reference: probability, site id, and the model's (seed, step counter)."""
    __slots__ = ("p", "site", "seed", "step")

    def __init__(self, p: float, site: int, seed: int, step: Optional[torch.Tensor]):
        self.p, self.site, self.seed, self.step = float(p), int(site), int(seed), step

    def c(self) -> L.Dropout:
        return L.make_dropout(self.p, self.site, self.seed, self.step)

    @property
    def on(self) -> bool:
        return self.p > 0.0


def _cdrop(drop: Optional[Drop]) -> L.Dropout:
    return drop.c() if (drop is not None and drop.on) else L.NO_DROPOUT


def _byref_drop(drop: Optional[Drop]):
    return C.byref(_cdrop(drop))


# ------------------------------------------------------------------------------------------------ raw GEMM launcher
def gemm(A, B, Cout, M, N, K, *, nseg=1, nbatch=1, lda, ldb, ldc, transA=False, transB=True, a_seg_off=0, b_seg_off=0,
         a_batch_off=0, b_batch_off=0, c_batch_off=0, shiftA: Sequence[int] = (), shiftB: Sequence[int] = (), T=0,
         bias=None, relu=False, drop: Optional[Drop] = None, residual=None, ldr=0, accumulate=False,
         a_ptr_off=0, b_ptr_off=0, c_ptr_off=0):
    """Launch pka_gemm_f32.  *_ptr_off are element offsets added to the base pointers (column blocks of packed buffers)."""
    L.require_cuda(A, B, Cout)
    if A.dtype != torch.float32 or B.dtype != torch.float32 or Cout.dtype != torch.float32:
        raise RuntimeError("pka_gemm_f32 needs fp32 operands")
    d = L.GemmDesc()
    d.A = A.data_ptr() + 4 * a_ptr_off
    d.B = B.data_ptr() + 4 * b_ptr_off
    d.C = Cout.data_ptr() + 4 * c_ptr_off
    d.bias = bias.data_ptr() if bias is not None else 0
    d.residual = residual.data_ptr() if residual is not None else 0
    d.M, d.N, d.K, d.nseg, d.nbatch = M, N, K, nseg, nbatch
    d.lda, d.ldb, d.ldc, d.ldr = lda, ldb, ldc, ldr
    d.transA, d.transB = int(transA), int(transB)
    d.a_seg_off, d.b_seg_off = a_seg_off, b_seg_off
    d.a_batch_off, d.b_batch_off, d.c_batch_off = a_batch_off, b_batch_off, c_batch_off
    for i, s in enumerate(shiftA):
        d.shiftA[i] = int(s)
    for i, s in enumerate(shiftB):
        d.shiftB[i] = int(s)
    d.T = T
    d.relu = int(relu)
    d.accumulate = int(accumulate)
    d.drop = _cdrop(drop)
    ws = None
    plain = bias is None and not relu and residual is None and (drop is None or not drop.on)
    if transA and nseg == 1 and plain and K >= 256:
        # weight gradients reduce over all frames but have few output tiles: split K so the grid fills the 148 SMs
        tiles = ((M + 63) // 64) * ((N + 63) // 64) * nbatch
        splitk = max(1, min((K + 63) // 64, (2 * 148 + tiles - 1) // tiles))
        if splitk > 1:
            ws = torch.empty(splitk * nbatch * M * N, device=A.device, dtype=torch.float32)
            d.splitk = splitk
            d.splitk_ws = ws.data_ptr()
    L.check(L.lib().pka_gemm_f32(C.byref(d), L.stream_ptr()), "gemm_f32")


# ------------------------------------------------------------------------------------------------ gradient slots
# FusedAdam keeps all gradients in one flat arena.  Backward kernels write a parameter's gradient straight into its
# arena slot: the autograd Function returns a fresh view of the slot, AccumulateGrad (p.grad is None after
# zero_grad) adopts that view instead of adding into a zeroed buffer -- no per-parameter `grad += g` kernels, no
# arena memset.  Without a FusedAdam (or when p.grad already holds something to accumulate into) a plain buffer is used.
_GRAD_SLOTS = {}
_SLOT_CLAIMS = set()          # arena slots handed out since the last forward pass (see grad_buffer)


def register_grad_slots(optimizer):
    import weakref
    ref = weakref.ref(optimizer)
    for i, p in enumerate(optimizer._train):
        _GRAD_SLOTS[p.data_ptr()] = (ref, i)


def grad_buffer(param_like: torch.Tensor, shape=None) -> torch.Tensor:
    """fp32 buffer for the gradient of the parameter `param_like` is (a view of): its arena slot when possible."""
    shape = tuple(shape) if shape is not None else tuple(param_like.shape)
    ent = _GRAD_SLOTS.get(param_like.data_ptr())
    if ent is not None:
        opt = ent[0]()
        if opt is not None:
            p = opt._train[ent[1]]
            if p.grad is None and p.data_ptr() == param_like.data_ptr() and p.numel() == param_like.numel():
                # a parameter that feeds two autograd nodes (tied weights, a module called twice) asks twice in one
                # backward pass: only the first request gets the slot, the second a plain buffer that autograd adds
                claim = (id(opt), ent[1])
                if claim not in _SLOT_CLAIMS:
                    _SLOT_CLAIMS.add(claim)
                    off = opt._offsets[ent[1]]
                    return opt.flat_grad[off:off + p.numel()].view(shape)
    return torch.empty(shape, device=param_like.device, dtype=torch.float32)


# ------------------------------------------------------------------------------------------------ deferred work
# The TIMIT-config step is launch bound.  Two kinds of backward work are NOT on the critical
# path of the backward pass -- nothing reads their result before the optimiser (or the gradient all-reduce) does:
#   * the fixed-order sums of split partials (weight-gradient splits, bias / LayerNorm column sums, packed per-head
#     gradients): round 1 ran 57 small "finish" kernels per step;
#   * the decoder-shaped weight gradients themselves (a few output tiles each, ~6 us per launch, 23 per step).
# Backward functions therefore only *record* them; one grouped weight-gradient launch and one reduction launch run at the
# end of the backward pass (autograd engine callback), or earlier when a data-parallel gradient bucket is about to leave.
# Summation orders are unchanged (bit-reproducible).  PKA_DEFER=0 restores one launch per reduction.
DEFER_ENABLED = os.environ.get("PKA_DEFER", "1") != "0"


class _Deferred:
    def __init__(self):
        self.lock = threading.Lock()
        self.wgrads, self.jobs, self.keep, self.queued = [], [], [], False
        self.side = {}               # device index -> side stream for work that is off the backward pass's critical path
        self.side_used = None        # the side stream that has work in flight since the last flush


_DEFER = _Deferred()
# Work nothing in the backward pass waits for (the full-wave TDNN weight-gradient GEMMs, the embedding scatter-add) runs
# on a side stream next to the data-gradient chain instead of in line with it; flush_deferred() joins it before the
# reductions.  Inside a CUDA-graph capture this becomes a fork / join of the graph.  PKA_SIDE=0 keeps one stream.
SIDE_ENABLED = os.environ.get("PKA_SIDE", "1") != "0"


class _SideStream:
    """`with _SideStream(*tensors_kept_alive):` -- the body's launches go to the side stream, ordered after everything
    already enqueued on the current stream.  Falls through to the current stream when no flush is scheduled."""

    def __init__(self, *keep):
        self.keep, self.ctx = keep, None

    def __enter__(self):
        if not (SIDE_ENABLED and _DEFER.queued):
            return self
        cur = torch.cuda.current_stream()
        side = _DEFER.side.get(cur.device_index)
        if side is None:
            side = _DEFER.side[cur.device_index] = torch.cuda.Stream(device=cur.device)
        side.wait_stream(cur)
        with _DEFER.lock:
            _DEFER.side_used = side
            _DEFER.keep.append(self.keep)         # operands stay allocated until the join
        self.ctx = torch.cuda.stream(side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _defer_schedule() -> bool:
    """True if work recorded now will be flushed by the running backward pass (callback installed); False when deferral
    is off or we are not inside a backward pass (the caller then finishes its reductions itself)."""
    if not DEFER_ENABLED:
        return False
    d = _DEFER
    if d.queued:
        return True
    try:
        torch.autograd.Variable._execution_engine.queue_callback(flush_deferred)
    except RuntimeError:
        return False
    d.queued = True
    return True


def defer_reduce(src: torch.Tensor, dst: torch.Tensor, n: int, splits: int, split_stride: int, *, src_off: int = 0,
                 kind: int = L.REDUCE_PLAIN, accumulate: bool = False, D: int = 0, dk: int = 0):
    """Record dst[...] (+)= sum_s src[src_off + s*split_stride + e] for pka_reduce_jobs; `src` is kept alive until the flush."""
    j = L.ReduceJob()
    j.src, j.dst = src.data_ptr() + 4 * src_off, dst.data_ptr()
    j.n, j.split_stride, j.splits, j.kind, j.accumulate, j.D, j.dk = n, split_stride, splits, kind, int(accumulate), D, dk
    # only the partials are kept alive: `dst` is a gradient tensor autograd is about to adopt, and an extra reference
    # would make AccumulateGrad clone it (before the sum has been written) instead of taking it over
    with _DEFER.lock:
        _DEFER.jobs.append(j)
        _DEFER.keep.append(src)


def flush_deferred():
    """Launch everything recorded so far: first the grouped weight gradients (their partials feed the sums), then the
    reductions.  Called by the autograd engine at the end of a backward pass, by parallel.GradAllReduce before a bucket
    leaves, and defensively by FusedAdam.step()."""
    d = _DEFER
    with d.lock:
        wg, jobs, keep, side = d.wgrads, d.jobs, d.keep, d.side_used
        d.wgrads, d.jobs, d.keep, d.queued, d.side_used = [], [], [], False, None
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)           # join: the side stream's partials feed the sums below
    if wg:
        arr = (L.TcDesc * len(wg))(*wg)
        L.check(L.lib().pka_gemm_tc_wgrad_group(arr, len(wg), L.stream_ptr()), "gemm_tc_wgrad_group")
    if jobs:
        arr = (L.ReduceJob * len(jobs))(*jobs)
        L.check(L.lib().pka_reduce_jobs(arr, len(jobs), L.stream_ptr()), "reduce_jobs")
    del keep


def reset_deferred():
    """Drop anything a failed backward pass may have left behind (called at the start of a forward pass)."""
    d = _DEFER
    with d.lock:
        d.wgrads, d.jobs, d.keep, d.queued, d.side_used = [], [], [], False, None
    _COLSUM_HINTS.clear()
    _PACKS.clear()
    _IN_DROP.clear()
    _SLOT_CLAIMS.clear()


# LayerNorm backward also leaves the column sums of the gradient it hands to the layer below (its dx) as partial rows in
# its workspace: when that layer is a bias-carrying linear map fed by exactly this tensor, its bias gradient is those
# sums -- no separate column-sum pass over dx.  Keyed by the address of dx; cleared at every forward pass.
_COLSUM_HINTS = {}


def colsum(x2d: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: Optional[bool] = None,
           defer: bool = False) -> torch.Tensor:
    rows, n = x2d.shape
    acc = (out is not None) if accumulate is None else bool(accumulate)
    if out is None:
        out = torch.empty(n, device=x2d.device, dtype=torch.float32)
    chunks = L.lib().pka_colsum_chunks(C.c_int64(rows))
    ws = torch.empty(chunks * n, device=x2d.device, dtype=torch.float32)
    deferred = defer and _defer_schedule()
    L.check(L.lib().pka_colsum(L.ptr(x2d), C.c_void_p(0) if deferred else L.ptr(out), L.ptr(ws), L.dtype_code(x2d),
                               C.c_int64(rows), n, x2d.stride(0), int(acc), L.stream_ptr()), "colsum")
    if deferred:
        parts = L.lib().pka_colsum_parts(L.ptr(x2d), L.ptr(ws), L.dtype_code(x2d), C.c_int64(rows), n, x2d.stride(0))
        defer_reduce(ws, out, n, parts, n, accumulate=acc)
    return out


# ------------------------------------------------------------------------------------------------ front-end
def frontend(feats: torch.Tensor, lengths: Optional[torch.Tensor], fold: int, ctx: Sequence[int], cmvn_mode: int = 0,
             out_dtype=torch.float32) -> torch.Tensor:
    """[B,T,F] fp32 -> [B, T//fold, len(ctx)*F*fold]: optional CMVN, frame folding, frame splicing in one pass."""
    L.require_cuda(feats, lengths)
    feats = feats.contiguous()
    B, T, F = feats.shape
    out = torch.empty(B, T // fold, len(ctx) * F * fold, device=feats.device, dtype=out_dtype)
    ctx_arr = (C.c_int32 * len(ctx))(*[int(c) for c in ctx])
    stats = torch.empty(B * 2 * F, device=feats.device, dtype=torch.float32) if cmvn_mode else None
    if lengths is not None:
        lengths = lengths.to(torch.int32).contiguous()
    L.check(L.lib().pka_frontend_fwd(L.ptr(feats), L.ptr(lengths), L.ptr(out), L.dtype_code(out), B, T, F, fold, ctx_arr,
                                     len(ctx), cmvn_mode, L.ptr(stats), L.stream_ptr()), "frontend_fwd")
    return out


# ------------------------------------------------------------------------------------------------ linear family
class _LinearFn(torch.autograd.Function):
    """y = dropout(relu(splice_ctx(x) @ W^T + b)): BottleLinear / TDNNLayer / Conv1d(k=1)."""

    @staticmethod
    def forward(ctx, x, weight, bias, splice, relu, drop, residual):
        L.require_cuda(x, weight, bias, residual)
        lead = x.shape[:-1]
        kin = x.shape[-1]
        x2 = x.reshape(-1, kin)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M = x2.shape[0]
        n_ctx = len(splice) if splice else 1
        w2 = weight.reshape(weight.shape[0], -1)
        N = w2.shape[0]
        assert w2.shape[1] == n_ctx * kin, "weight [%d,%d] does not match %d contexts x %d" % (N, w2.shape[1], n_ctx, kin)
        T = x.shape[-2] if splice else 0
        y = torch.empty(M, N, device=x.device, dtype=torch.float32)
        r2 = None
        if residual is not None:
            assert not relu, "residual is added after dropout; combining it with ReLU is not a reference pattern"
            r2 = residual.reshape(M, N).contiguous()
        gemm(x2, w2, y, M, N, kin, nseg=n_ctx, lda=kin, ldb=n_ctx * kin, ldc=N, transB=True, b_seg_off=kin,
             shiftA=splice or (), T=T, bias=bias, relu=relu, drop=drop, residual=r2, ldr=N)
        ctx.save_for_backward(x2, w2, y if relu else None, bias)
        ctx.meta = (lead, kin, n_ctx, tuple(splice) if splice else (), T, relu, drop, bias is not None, weight.shape,
                    residual is not None)
        return y.view(*lead, N)

    @staticmethod
    def backward(ctx, dy):
        x2, w2, y, bias = ctx.saved_tensors
        lead, kin, n_ctx, splice, T, relu, drop, has_bias, wshape, has_res = ctx.meta
        M, N = x2.shape[0], w2.shape[0]
        dy2 = dy.reshape(M, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        if relu:
            dz = torch.empty_like(dy2)
            scale = 1.0 / (1.0 - drop.p) if (drop is not None and drop.on) else 1.0
            L.check(L.lib().pka_relu_drop_bwd(L.ptr(dy2), L.ptr(y), L.ptr(dz), L.PKA_F32, C.c_int64(M * N), C.c_float(scale),
                                              L.stream_ptr()), "relu_drop_bwd")
        elif drop is not None and drop.on:
            dz = torch.empty_like(dy2)
            L.check(L.lib().pka_dropout_bwd(L.ptr(dy2), L.ptr(dz), L.PKA_F32, C.c_int64(M * N), _byref_drop(drop),
                                            L.stream_ptr()), "dropout_bwd")
        else:
            dz = dy2
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, kin, device=dy.device, dtype=torch.float32)
            # dx[m,i] = sum_s sum_o dz[m - ctx_s, o] * W[o, s*kin + i]
            gemm(dz, w2, dx, M, kin, N, nseg=n_ctx, lda=N, ldb=n_ctx * kin, ldc=kin, transB=False, b_seg_off=kin,
                 shiftA=[-c for c in splice], T=T)
            dx = dx.view(*lead, kin)
        if ctx.needs_input_grad[1]:
            dw = grad_buffer(w2, (N, n_ctx * kin))
            # dW[o, s*kin + i] = sum_m dz[m,o] * x[m + ctx_s, i]   (batch over contexts, frame shift on the reduction index)
            gemm(dz, x2, dw, N, kin, M, nbatch=n_ctx, lda=N, ldb=kin, ldc=n_ctx * kin, transA=True, transB=False,
                 c_batch_off=kin, shiftB=splice, T=T)
            dw = dw.view(wshape)
        if has_bias and ctx.needs_input_grad[2]:
            db = colsum(dz, out=grad_buffer(bias), accumulate=False)
        dres = dy if (has_res and ctx.needs_input_grad[6]) else None
        return dx, dw, db, None, None, None, dres


def linear(x, weight, bias=None, splice: Optional[Sequence[int]] = None, relu: bool = False, drop: Optional[Drop] = None,
           residual=None):
    """dropout(relu(splice(x) @ W^T + b)) [+ residual]."""
    return _LinearFn.apply(x, weight, bias, list(splice) if splice else None, relu, drop, residual)


def affine_kn(x, weight_kn, bias=None):
    """x @ W + b for a frozen [in, out] matrix (LDALayer, L/pytorch/TDNN.py:53-55).  No autograd: the reference keeps
    these parameters requires_grad=False and the input is the feature tensor."""
    L.require_cuda(x, weight_kn, bias)
    lead, kin = x.shape[:-1], x.shape[-1]
    x2 = x.detach().reshape(-1, kin)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    w = weight_kn.detach()
    if not w.is_contiguous():
        w = w.contiguous()
    assert w.shape[0] == kin
    N = w.shape[1]
    y = torch.empty(x2.shape[0], N, device=x.device, dtype=torch.float32)
    gemm(x2, w, y, x2.shape[0], N, kin, lda=kin, ldb=N, ldc=N, transB=False, bias=bias.detach() if bias is not None else None)
    return y.view(*lead, N)


class _HeadProjFn(torch.autograd.Function):
    """Packed per-head projections: out[..., p*H*dk + h*dk + j] = sum_d x[..., d] * w_p[h, d, j] for p in 0..len(ws)-1.
    Replaces q.repeat(n_head)+bmm (T/SubLayers.py:49-56): no replication, heads are column blocks."""

    @staticmethod
    def forward(ctx, x, link, *ws):
        L.require_cuda(x, *ws)
        if x.dtype == torch.bfloat16:
            return _HeadProjFn._forward_tc(ctx, x, ws, link)
        if link is not None:
            link.armed = False
        ctx.tc = False
        lead, D = x.shape[:-1], x.shape[-1]
        x2 = x.reshape(-1, D)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M = x2.shape[0]
        H, _, dk = ws[0].shape
        P = len(ws)
        out = torch.empty(M, P * H * dk, device=x.device, dtype=torch.float32)
        for p, w in enumerate(ws):
            assert w.shape == (H, D, dk) and w.is_contiguous()
            gemm(x2, w, out, M, dk, D, nbatch=H, lda=D, ldb=dk, ldc=P * H * dk, transB=False, b_batch_off=D * dk,
                 c_batch_off=dk, c_ptr_off=p * H * dk)
        ctx.save_for_backward(x2, *ws)
        ctx.meta = (lead, D, H, dk, P)
        return out.view(*lead, P * H * dk)

    @staticmethod
    def _forward_tc(ctx, x, ws, link=None):
        """bf16 x [Bt,T,D]: one tcgen05 GEMM for all P*H heads; the packed q|k|v (or k|v) buffer stays bf16 because its
        only consumer is the tensor-core attention kernel."""
        assert x.dim() == 3 and x.is_contiguous()
        Bt, T, D = x.shape
        H, _, dk = ws[0].shape
        P = len(ws)
        ntot = P * H * dk
        needs_dx = ctx.needs_input_grad[0]
        if OPERANDS is not None and all(w.is_contiguous() for w in ws):
            wf, wd = OPERANDS.heads(ws)                             # refreshed once per forward pass, not per op
        else:
            wf = torch.empty(ntot, D, device=x.device, dtype=torch.bfloat16)
            wd = torch.empty(D, ntot, device=x.device, dtype=torch.bfloat16) if needs_dx else None
            wp = [L.ptr(w.detach()) for w in ws] + [C.c_void_p(0)] * (3 - P)
            L.check(L.lib().pka_head_weight_relayout(wp[0], wp[1], wp[2], P, H, D, dk, L.ptr(wf), L.ptr(wd), L.stream_ptr()),
                    "head_weight_relayout")
        if link is not None:
            link.armed = bool(needs_dx)
        ctx.link = link if (link is not None and link.armed) else None
        out = gemm_tc_rows(x, wf, Bt, T, ntot, D, lda=D, ldb=D, out_dtype=torch.bfloat16)
        ctx.tc = True
        ctx.save_for_backward(x, wd, *ws)
        ctx.meta = (Bt, T, D, H, dk, P)
        return out

    @staticmethod
    def _backward_tc(ctx, dy):
        x, wd, *ws = ctx.saved_tensors
        Bt, T, D, H, dk, P = ctx.meta
        ntot = P * H * dk
        dz = dy if (dy.dtype == torch.bfloat16 and dy.is_contiguous()) else gate_to_bf16(dy, Bt, T, ntot)
        dx = None
        addend = ctx.link.take() if ctx.link is not None else None       # residual-branch gradient of the sub-layer input
        if ctx.needs_input_grad[0]:
            dx = gemm_tc_rows(dz, wd, Bt, T, D, ntot, lda=ntot, ldb=ntot, addend=addend)
        dws = [None] * P
        if any(ctx.needs_input_grad[2:2 + P]):
            dws = [grad_buffer(ws[p]) if ctx.needs_input_grad[2 + p] else None for p in range(P)]
            gemm_tc_wgrad(dz, x, Bt, T, ntot, D, 1, (0,), defer=True, heads=(dws, P, H, D, dk))
        return (dx, None, *dws)

    @staticmethod
    def backward(ctx, dy):
        if ctx.tc:
            return _HeadProjFn._backward_tc(ctx, dy)
        x2, *ws = ctx.saved_tensors
        lead, D, H, dk, P = ctx.meta
        M = x2.shape[0]
        dy2 = dy.reshape(M, P * H * dk)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, D, device=dy.device, dtype=torch.float32)
            for p, w in enumerate(ws):
                # dx[m,d] (+)= sum_h sum_j dy[m, p*H*dk + h*dk + j] * w_p[h,d,j]
                gemm(dy2, w, dx, M, D, dk, nseg=H, lda=P * H * dk, ldb=dk, ldc=D, transB=True, a_seg_off=dk,
                     b_seg_off=D * dk, a_ptr_off=p * H * dk, accumulate=(p > 0))
            dx = dx.view(*lead, D)
        dws = []
        for p, w in enumerate(ws):
            if not ctx.needs_input_grad[2 + p]:
                dws.append(None)
                continue
            dw = grad_buffer(w)
            # dw_p[h,d,j] = sum_m x[m,d] * dy[m, p*H*dk + h*dk + j]
            gemm(x2, dy2, dw, D, dk, M, nbatch=H, lda=D, ldb=P * H * dk, ldc=dk, transA=True, transB=False,
                 b_batch_off=dk, c_batch_off=D * dk, b_ptr_off=p * H * dk)
            dws.append(dw)
        return (dx, None, *dws)


def head_proj(x, *ws, link: Optional[ResidualLink] = None):
    return _HeadProjFn.apply(x, link, *ws)


# ------------------------------------------------------------------------------------------------ attention
class _AttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qbuf, kvbuf, key_mask, H, dk, band, scale, drop, want_probs):
        """self-attention: qbuf = packed [B,L,3*H*dk] (q|k|v), kvbuf None.
        cross-attention: qbuf = [B,Lq,H*dk], kvbuf = packed [B,Lk,2*H*dk] (k|v)."""
        L.require_cuda(qbuf, kvbuf, key_mask)
        ctx.set_materialize_grads(False)          # no zero-filled gradients for the statistics outputs
        qbuf = qbuf.contiguous()
        HD = H * dk
        B, Lq = qbuf.shape[0], qbuf.shape[1]
        if kvbuf is None:
            Lk, ldq = Lq, 3 * HD
            q_p, k_p, v_p, ldk = qbuf.data_ptr(), qbuf.data_ptr() + 4 * HD, qbuf.data_ptr() + 8 * HD, 3 * HD
        else:
            kvbuf = kvbuf.contiguous()
            Lk, ldq, ldk = kvbuf.shape[1], HD, 2 * HD
            q_p, k_p, v_p = qbuf.data_ptr(), kvbuf.data_ptr(), kvbuf.data_ptr() + 4 * HD
        key_mask = key_mask.to(torch.uint8).contiguous()
        assert key_mask.shape == (B, Lk)
        d = L.AttnDesc()
        d.B, d.H, d.Lq, d.Lk, d.dk, d.dv = B, H, Lq, Lk, dk, dk
        d.ldq, d.ldk, d.ldv, d.ldo = ldq, ldk, ldk, HD
        d.use_band = int(band is not None)
        d.band_start, d.band_end = (int(band[0]), int(band[1])) if band is not None else (0, 0)
        d.scale = float(scale)
        d.drop = _cdrop(drop)
        out = torch.empty(B, Lq, HD, device=qbuf.device, dtype=torch.float32)
        lse = torch.empty(B, H, Lq, device=qbuf.device, dtype=torch.float32)
        probs = torch.empty(B, H, Lq, Lk, device=qbuf.device, dtype=torch.float32) if want_probs else None
        L.check(L.lib().pka_attn_fwd(C.byref(d), L.PKA_F32, C.c_void_p(q_p), C.c_void_p(k_p), C.c_void_p(v_p),
                                     L.ptr(key_mask), L.ptr(out), L.ptr(lse), L.ptr(probs), L.stream_ptr()), "attn_fwd")
        ctx.save_for_backward(qbuf, kvbuf, key_mask, out, lse)
        ctx.desc = d
        ctx.mark_non_differentiable(lse)
        if probs is not None:
            ctx.mark_non_differentiable(probs)
        return out, lse, probs

    @staticmethod
    def backward(ctx, dout, _dlse, _dprobs):
        if dout is None:
            return (None,) * 9
        qbuf, kvbuf, key_mask, out, lse = ctx.saved_tensors
        d = ctx.desc
        HD = d.H * d.dk
        dout = dout.contiguous()
        dqbuf = torch.empty_like(qbuf)
        if kvbuf is None:
            dkvbuf = None
            q_p, k_p, v_p = qbuf.data_ptr(), qbuf.data_ptr() + 4 * HD, qbuf.data_ptr() + 8 * HD
            dq_p, dk_p, dv_p = dqbuf.data_ptr(), dqbuf.data_ptr() + 4 * HD, dqbuf.data_ptr() + 8 * HD
        else:
            dkvbuf = torch.empty_like(kvbuf)
            q_p, k_p, v_p = qbuf.data_ptr(), kvbuf.data_ptr(), kvbuf.data_ptr() + 4 * HD
            dq_p, dk_p, dv_p = dqbuf.data_ptr(), dkvbuf.data_ptr(), dkvbuf.data_ptr() + 4 * HD
        delta = torch.empty_like(lse)
        L.check(L.lib().pka_attn_bwd(C.byref(d), L.PKA_F32, C.c_void_p(q_p), C.c_void_p(k_p), C.c_void_p(v_p),
                                     L.ptr(key_mask), L.ptr(out), L.ptr(dout), L.ptr(lse), L.ptr(delta),
                                     C.c_void_p(dq_p), C.c_void_p(dk_p), C.c_void_p(dv_p), L.stream_ptr()), "attn_bwd")
        return dqbuf, dkvbuf, None, None, None, None, None, None, None


class _AttnTcFn(torch.autograd.Function):
    """bf16 attention on the tensor cores (pka_attn_tc_fwd: tcgen05 S = Q K^T / O = P V, TMA-fed, online softmax).
    Same packed-buffer convention as _AttnFn; q/k/v and the output are bf16, statistics fp32."""

    @staticmethod
    def forward(ctx, qbuf, kvbuf, key_mask, H, dk, band, scale, drop, out_fp32, kv_pack=None):
        L.require_cuda(qbuf, kvbuf, key_mask)
        ctx.set_materialize_grads(False)          # no zero-filled gradient for the lse output
        assert qbuf.dtype == torch.bfloat16 and (kvbuf is None or kvbuf.dtype == torch.bfloat16)
        qbuf = qbuf.contiguous()
        HD = H * dk
        B, Lq = qbuf.shape[0], qbuf.shape[1]
        ctx.kv_pack = None
        if kvbuf is None:
            Lk, ldq = Lq, 3 * HD
            q_p, k_p, v_p, ldk = qbuf.data_ptr(), qbuf.data_ptr() + 2 * HD, qbuf.data_ptr() + 4 * HD, 3 * HD
        else:
            # k|v may be a column block of a wider packed buffer (cross-attention K/V of all layers from one GEMM)
            if not (kvbuf.stride(2) == 1 and kvbuf.stride(0) == kvbuf.shape[1] * kvbuf.stride(1) and kvbuf.stride(1) % 8 == 0):
                kvbuf = kvbuf.contiguous()
            elif kv_pack is not None:
                ctx.kv_pack = kv_pack
            Lk, ldq, ldk = kvbuf.shape[1], HD, kvbuf.stride(1)
            q_p, k_p, v_p = qbuf.data_ptr(), kvbuf.data_ptr(), kvbuf.data_ptr() + 2 * HD
        key_mask = key_mask.to(torch.uint8).contiguous()
        assert key_mask.shape == (B, Lk)
        d = L.AttnDesc()
        d.B, d.H, d.Lq, d.Lk, d.dk, d.dv = B, H, Lq, Lk, dk, dk
        d.ldq, d.ldk, d.ldv, d.ldo = ldq, ldk, ldk, HD
        d.use_band = int(band is not None)
        d.band_start, d.band_end = (int(band[0]), int(band[1])) if band is not None else (0, 0)
        d.scale = float(scale)
        d.drop = _cdrop(drop)
        out = torch.empty(B, Lq, HD, device=qbuf.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
        lse = torch.empty(B, H, Lq, device=qbuf.device, dtype=torch.float32)
        L.check(L.lib().pka_attn_tc_fwd(C.byref(d), C.c_void_p(q_p), C.c_void_p(k_p), C.c_void_p(v_p), L.ptr(key_mask),
                                        L.ptr(out), L.dtype_code(out), L.ptr(lse), L.stream_ptr()), "attn_tc_fwd")
        ctx.save_for_backward(qbuf, kvbuf, key_mask, out, lse)
        ctx.desc = d
        ctx.mark_non_differentiable(lse)
        return out, lse

    @staticmethod
    def backward(ctx, dout, _dlse):
        if dout is None:
            return (None,) * 10
        qbuf, kvbuf, key_mask, out, lse = ctx.saved_tensors
        d = ctx.desc
        HD = d.H * d.dk
        dout = cast(dout, torch.bfloat16).contiguous()
        out = cast(out, torch.bfloat16)
        dqbuf = torch.empty_like(qbuf)
        if kvbuf is None:
            dkvbuf = None
            q_p, k_p, v_p = qbuf.data_ptr(), qbuf.data_ptr() + 2 * HD, qbuf.data_ptr() + 4 * HD
            dq_p, dk_p, dv_p = dqbuf.data_ptr(), dqbuf.data_ptr() + 2 * HD, dqbuf.data_ptr() + 4 * HD
        else:
            if ctx.kv_pack is not None:            # gradient goes straight into its column block of the shared packed buffer
                dkvbuf = ctx.kv_pack.slot_like(kvbuf)
            elif kvbuf.is_contiguous():
                dkvbuf = torch.empty_like(kvbuf)
            else:
                dkvbuf = torch.empty_strided(kvbuf.shape, kvbuf.stride(), device=kvbuf.device, dtype=kvbuf.dtype)
            q_p, k_p, v_p = qbuf.data_ptr(), kvbuf.data_ptr(), kvbuf.data_ptr() + 2 * HD
            dq_p, dk_p, dv_p = dqbuf.data_ptr(), dkvbuf.data_ptr(), dkvbuf.data_ptr() + 2 * HD
        delta = torch.empty_like(lse)
        if ATTN_BWD_TC:
            L.check(L.lib().pka_attn_tc_bwd(C.byref(d), C.c_void_p(q_p), C.c_void_p(k_p), C.c_void_p(v_p), L.ptr(key_mask),
                                            L.ptr(out), L.ptr(dout), L.ptr(lse), L.ptr(delta), C.c_void_p(dq_p),
                                            C.c_void_p(dk_p), C.c_void_p(dv_p), L.stream_ptr()), "attn_tc_bwd")
        else:
            L.check(L.lib().pka_attn_bwd(C.byref(d), L.PKA_BF16, C.c_void_p(q_p), C.c_void_p(k_p), C.c_void_p(v_p),
                                         L.ptr(key_mask), L.ptr(out), L.ptr(dout), L.ptr(lse), L.ptr(delta),
                                         C.c_void_p(dq_p), C.c_void_p(dk_p), C.c_void_p(dv_p), L.stream_ptr()), "attn_bwd")
        return dqbuf, dkvbuf, None, None, None, None, None, None, None, None


ATTN_BWD_TC = True      # False: the SIMT flash backward on the same bf16 buffers (kept for A/B comparison in the tests)


def attention(qbuf, kvbuf, key_mask, n_head: int, d_k: int, band, scale: float, drop: Optional[Drop] = None,
              want_probs: bool = False):
    out, _lse, probs = _AttnFn.apply(qbuf, kvbuf, key_mask, n_head, d_k, band, scale, drop, want_probs)
    return out, probs


def attention_tc(qbuf, kvbuf, key_mask, n_head: int, d_k: int, band, scale: float, drop: Optional[Drop] = None,
                 out_fp32: bool = False):
    """bf16 packed projections -> bf16 (or fp32) context on the tcgen05 attention kernel; returns (out, lse)."""
    return _AttnTcFn.apply(qbuf, kvbuf, key_mask, n_head, d_k, band, scale, drop, out_fp32,
                           _PACKS.get(kvbuf.data_ptr()) if kvbuf is not None else None)


_PACKS = {}           # address of a column-block view -> GradPack of its base (cleared at every forward pass)


class _DropRec:
    """A dropout whose output feeds exactly one tensor-core linear layer: that layer applies the keep mask in the epilogue
    of its data-gradient GEMM (the mask is a pure function of the element index, and dX has the index space of the
    dropout's output), so the dropout's own backward kernel disappears.  `fused` is set by the consuming layer."""
    __slots__ = ("drop", "fused")

    def __init__(self, drop):
        self.drop, self.fused = drop, False


_IN_DROP = {}         # address of a dropout output with a single linear consumer -> _DropRec (cleared at every forward pass)


class GradPack:
    """Gradient side of a packed multi-consumer buffer: every consumer writes its gradient into its own column block of
    ONE buffer shaped like the forward buffer, so the producer's backward sees a single packed gradient."""

    def __init__(self, base: torch.Tensor):
        self.base_ptr, self.shape, self.stride, self.buf = base.data_ptr(), tuple(base.shape), base.stride(), None
        self.item, self.device, self.dtype = base.element_size(), base.device, base.dtype

    def slot_like(self, view: torch.Tensor) -> torch.Tensor:
        if self.buf is None:
            self.buf = torch.empty(self.shape, device=self.device, dtype=self.dtype)
        off = (view.data_ptr() - self.base_ptr) // self.item           # element offset of the view inside the base
        return self.buf.as_strided(view.shape, view.stride(), off)


class _CrossKVFn(torch.autograd.Function):
    """k|v projections of the encoder memory for ALL decoder layers in one tensor-core GEMM (the reference projects them
    per layer, T/SubLayers.py:49-56 inside each DecoderLayer's enc_attn): out[..., (l*2 + p)*H*dk + h*dk + j] with p = 0
    for w_ks, 1 for w_vs.  Returns one column-block view per layer; their gradients come back as column blocks of one
    packed buffer (GradPack), so backward is ONE data-gradient GEMM (K = n_layers*2*H*dk) and ONE weight-gradient."""

    @staticmethod
    def forward(ctx, enc, *ws):
        L.require_cuda(enc, *ws)
        assert enc.dtype == torch.bfloat16 and enc.dim() == 3 and enc.is_contiguous() and len(ws) % 2 == 0
        Bt, T, D = enc.shape
        H, _, dk = ws[0].shape
        P = len(ws)
        ntot = P * H * dk
        if OPERANDS is not None and all(w.is_contiguous() for w in ws):
            wf, wd = OPERANDS.heads(ws)
        else:
            cache = OperandCache()
            wf, wd = cache.heads(ws)
        out = gemm_tc_rows(enc, wf, Bt, T, ntot, D, lda=D, ldb=D, out_dtype=torch.bfloat16)
        ctx.save_for_backward(enc, wd, *ws)
        ctx.meta = (Bt, T, D, H, dk, P)
        ctx.pack = pack = GradPack(out)
        views = tuple(out[:, :, l * 2 * H * dk:(l + 1) * 2 * H * dk] for l in range(P // 2))
        for v in views[1:]:                        # (the first view shares the base address: registered last, below)
            _PACKS[v.data_ptr()] = pack
        _PACKS[views[0].data_ptr()] = pack
        return views

    @staticmethod
    def backward(ctx, *ds):
        enc, wd, *ws = ctx.saved_tensors
        Bt, T, D, H, dk, P = ctx.meta
        ntot, blk = P * H * dk, 2 * H * dk
        pack = ctx.pack
        ok = pack.buf is not None and all(
            d is not None and d.data_ptr() == pack.buf.data_ptr() + 2 * l * blk and d.stride(1) == ntot for l, d in enumerate(ds))
        if ok:
            dz = pack.buf
        else:                                      # a consumer bypassed the pack (or got no gradient): assemble it
            dz = torch.cat([d if d is not None else torch.zeros(Bt, T, blk, device=enc.device, dtype=torch.bfloat16) for d in ds], dim=2)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = gemm_tc_rows(dz, wd, Bt, T, D, ntot, lda=ntot, ldb=ntot)
        dws = [None] * P
        if any(ctx.needs_input_grad[1:1 + P]):
            dws = [grad_buffer(ws[p]) if ctx.needs_input_grad[1 + p] else None for p in range(P)]
            # the packed weight gradient [(l,p,h,j), d] is summed per tensor into the reference layout [H, D, dk]
            gemm_tc_wgrad(dz, enc, Bt, T, ntot, D, 1, (0,), defer=True, heads=(dws, P, H, D, dk))
        return (dx, *dws)


def cross_kv_proj(enc, layer_weights):
    """[(w_ks, w_vs) per layer] -> list of bf16 k|v buffers [B, T, 2*H*dk], one per layer (column blocks of one GEMM output)."""
    flat = [w for pair in layer_weights for w in pair]
    return list(_CrossKVFn.apply(enc, *flat))


# ------------------------------------------------------------------------------------------------ add + LayerNorm
class ResidualLink:
    """Side channel between the two backward functions of a post-LN sub-layer  y = LN(f(x) + x).

    Autograd would add the two gradient branches of x (through f and through the residual) with an element-wise kernel
    of its own.  Here the LayerNorm backward (which runs first) hands its residual-branch gradient over through this
    object instead of returning it, and the FIRST op of f (a tensor-core projection) adds it in the epilogue of its
    data-gradient GEMM: dX = dZ.W + dResidual, one launch and one pass less per sub-layer.  `armed` is set by the
    consuming op in its forward when it will really produce dX on the tensor-core path."""
    __slots__ = ("armed", "dres")

    def __init__(self):
        self.armed, self.dres = False, None

    def take(self):
        d, self.dres = self.dres, None
        return d


class _AddLayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, a, b, eps, drop, link=None):
        L.require_cuda(x, residual, a, b)
        shape = x.shape
        D = shape[-1]
        x2 = x.reshape(-1, D).contiguous()
        r2 = residual.reshape(-1, D).contiguous() if residual is not None else None
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rinv = torch.empty(rows, device=x.device, dtype=torch.float32)
        L.check(L.lib().pka_add_layernorm_fwd(L.ptr(x2), L.ptr(r2), L.ptr(a), L.ptr(b), L.ptr(y), L.ptr(mean), L.ptr(rinv),
                                              L.dtype_code(x2), rows, D, C.c_float(eps), _byref_drop(drop), L.stream_ptr()),
                "add_layernorm_fwd")
        ctx.save_for_backward(x2, r2, a, mean, rinv, b)
        ctx.meta = (shape, eps, drop)
        ctx.link = link if (link is not None and link.armed and r2 is not None) else None
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        x2, r2, a, mean, rinv, b_par = ctx.saved_tensors
        shape, eps, drop = ctx.meta
        rows, D = x2.shape
        dy2 = dy.reshape(rows, D).contiguous()
        dres = torch.empty_like(x2)
        use_drop = drop is not None and drop.on
        dx = torch.empty_like(x2) if use_drop else None
        da, db = grad_buffer(a), grad_buffer(b_par)                 # overwritten by the kernel
        nblk = L.lib().pka_ln_bwd_blocks(rows)
        ws = torch.empty(3 * D * nblk, device=dy.device, dtype=torch.float32)    # per CTA: [da | db | column sums of dx]
        deferred = _defer_schedule()                                # the gain / offset partial rows are summed at the flush
        null = C.c_void_p(0)
        L.check(L.lib().pka_add_layernorm_bwd(L.ptr(dy2), L.ptr(x2), L.ptr(r2), L.ptr(a), L.ptr(mean), L.ptr(rinv), L.ptr(dx),
                                              L.ptr(dres), null if deferred else L.ptr(da), null if deferred else L.ptr(db),
                                              L.ptr(ws), L.dtype_code(x2), rows, D,
                                              C.c_float(eps), _byref_drop(drop), L.stream_ptr()), "add_layernorm_bwd")
        if deferred:
            defer_reduce(ws, da, D, nblk, 3 * D)
            defer_reduce(ws, db, D, nblk, 3 * D, src_off=D)
        dres_v = dres.view(shape)
        dx_v = dx.view(shape) if use_drop else dres_v
        if deferred:
            _COLSUM_HINTS[dx_v.data_ptr()] = (ws, nblk, rows, D)
        if ctx.link is not None:                  # the residual branch travels to the sub-layer's first GEMM (see ResidualLink)
            if use_drop:
                ctx.link.dres = dres_v
                return dx_v, None, da, db, None, None, None
            # without dropout dx and dres are ONE buffer: it must stay intact for the projection's backward, so the
            # consumer reads it as the addend and this branch returns it as is
            ctx.link.dres = dres_v
            return dx_v, None, da, db, None, None, None
        return dx_v, (dres_v if r2 is not None else None), da, db, None, None, None


def add_layer_norm(x, residual, a, b, eps: float = 1e-3, drop: Optional[Drop] = None, link: Optional[ResidualLink] = None):
    """LayerNormalization(dropout(x) + residual) with the reference's formula (T/Modules.py:42-51)."""
    return _AddLayerNormFn.apply(x, residual, a, b, eps, drop, link)


# ------------------------------------------------------------------------------------------------ embedding, positions
class _EmbedPosFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tok, emb, pos, drop, padding_idx, out_dtype):
        L.require_cuda(tok, emb, pos)
        tok = tok.to(torch.int64).contiguous()
        B, Ln = tok.shape
        V, D = emb.shape
        assert pos.shape[0] >= Ln, "sequence length %d exceeds the position table (%d)" % (Ln, pos.shape[0])
        out = torch.empty(B, Ln, D, device=emb.device, dtype=out_dtype)
        L.check(L.lib().pka_embed_pos_fwd(L.ptr(tok), L.ptr(emb), L.ptr(pos), L.ptr(out), L.dtype_code(out), B, Ln, D, V,
                                          _byref_drop(drop), L.stream_ptr()), "embed_pos_fwd")
        ctx.save_for_backward(tok, emb)
        ctx.meta = (V, D, drop, padding_idx)
        return out

    @staticmethod
    def backward(ctx, dout):
        tok, emb = ctx.saved_tensors
        V, D, drop, padding_idx = ctx.meta
        B, Ln = tok.shape
        dout = dout.contiguous()
        demb = grad_buffer(emb)                                      # every row is written by the kernel
        _defer_schedule()
        with _SideStream(tok, dout):              # nothing downstream in the backward pass reads the embedding gradient
            L.check(L.lib().pka_embed_bwd(L.ptr(tok), L.ptr(dout), L.ptr(demb), L.dtype_code(dout), B, Ln, D, V, padding_idx,
                                          _byref_drop(drop), L.stream_ptr()), "embed_bwd")
        return None, demb, None, None, None, None


def embed_pos(tok, emb, pos, drop: Optional[Drop] = None, padding_idx: int = 0, out_dtype=torch.float32):
    return _EmbedPosFn.apply(tok, emb, pos, drop, padding_idx, out_dtype)


class _CastFn(torch.autograd.Function):
    """dtype conversion between the bf16 activation stream and fp32 consumers (pka_cast); the gradient is cast back."""

    @staticmethod
    def forward(ctx, x, dtype):
        L.require_cuda(x)
        x = x.contiguous()
        out = torch.empty(x.shape, device=x.device, dtype=dtype)
        L.check(L.lib().pka_cast(L.ptr(x), L.dtype_code(x), L.ptr(out), L.dtype_code(out), C.c_int64(x.numel()), L.stream_ptr()),
                "cast")
        ctx.src_dtype = x.dtype
        return out

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        dx = torch.empty(dy.shape, device=dy.device, dtype=ctx.src_dtype)
        L.check(L.lib().pka_cast(L.ptr(dy), L.dtype_code(dy), L.ptr(dx), L.dtype_code(dx), C.c_int64(dy.numel()), L.stream_ptr()),
                "cast")
        return dx, None


def cast(x, dtype):
    return x if x.dtype == dtype else _CastFn.apply(x, dtype)


class _AddRowvecDropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rowvec, period, drop, single_consumer=False):
        L.require_cuda(x, rowvec)
        x = x.contiguous()
        D = x.shape[-1]
        rows = x.numel() // D
        if rowvec is not None:
            assert rowvec.shape[0] >= period and rowvec.shape[1] == D and rowvec.is_contiguous()
        out = torch.empty_like(x)
        L.check(L.lib().pka_add_rowvec_dropout_fwd(L.ptr(x), L.ptr(rowvec), L.ptr(out), L.dtype_code(x), C.c_int64(rows), D,
                                                   period, _byref_drop(drop), L.stream_ptr()), "add_rowvec_dropout_fwd")
        ctx.drop = drop
        ctx.rec = None
        if single_consumer and drop is not None and drop.on and out.dtype == torch.bfloat16:
            ctx.rec = _IN_DROP[out.data_ptr()] = _DropRec(drop)
        return out

    @staticmethod
    def backward(ctx, dy):
        drop = ctx.drop
        if drop is None or not drop.on or (ctx.rec is not None and ctx.rec.fused):
            return dy, None, None, None, None     # (fused: the consuming layer's data-gradient GEMM applied the mask)
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        L.check(L.lib().pka_dropout_bwd(L.ptr(dy), L.ptr(dx), L.dtype_code(dy), C.c_int64(dy.numel()), _byref_drop(drop),
                                        L.stream_ptr()), "dropout_bwd")
        return dx, None, None, None, None


def add_pos_dropout(x, pos_table: Optional[torch.Tensor], drop: Optional[Drop] = None, single_consumer: bool = False):
    """dropout(x + pos_table[arange(x.size(1))]) (T/Models.py:164-165); pos_table None = plain dropout.
    `single_consumer`: the result feeds exactly one tensor-core linear layer, which may then take over the mask of the
    backward pass (see _DropRec)."""
    if pos_table is None and (drop is None or not drop.on):
        return x
    period = x.shape[-2]
    if pos_table is not None:
        assert pos_table.shape[0] >= period, "sequence length %d exceeds the position table (%d)" % (period, pos_table.shape[0])
    return _AddRowvecDropoutFn.apply(x, pos_table, period, drop, single_consumer)


# ------------------------------------------------------------------------------------------------ loss
class _CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits2d, goal, smoothing, eps):
        out3 = _CE_OUT.pop() if _CE_OUT else None          # caller-owned result buffer (not an autograd input)
        L.require_cuda(logits2d, goal, out3)
        ctx.set_materialize_grads(False)
        logits2d = logits2d.contiguous()
        goal = goal.to(torch.int64).contiguous()
        N, V = logits2d.shape
        if out3 is None:
            out3 = torch.empty(3, device=logits2d.device, dtype=torch.float32)
        else:
            assert out3.dtype == torch.float32 and out3.numel() == 3 and out3.is_contiguous()
        lse = torch.empty(N, device=logits2d.device, dtype=torch.float32)
        nblk = L.lib().pka_ce_blocks(N)
        ws = torch.empty(3 * nblk, device=logits2d.device, dtype=torch.float32)
        done = _CE_DONE.get(logits2d.device.index)
        if done is None:                          # zero once; the kernel's last CTA resets it after summing the partials
            done = _CE_DONE[logits2d.device.index] = torch.zeros(1, device=logits2d.device, dtype=torch.int32)
        L.check(L.lib().pka_ce_fwd_fused(L.ptr(logits2d), L.ptr(goal), L.dtype_code(logits2d), N, V, int(smoothing),
                                         C.c_float(eps), L.ptr(out3), L.ptr(lse), L.ptr(ws), L.ptr(done), L.stream_ptr()),
                "ce_fwd")
        ctx.save_for_backward(logits2d, goal, lse)
        ctx.meta = (int(smoothing), eps)
        loss = out3[0]
        stats = out3[1:]
        ctx.mark_non_differentiable(stats)
        return loss, stats

    @staticmethod
    def backward(ctx, gloss, _gstats):
        if gloss is None:
            return None, None, None, None
        logits2d, goal, lse = ctx.saved_tensors
        smoothing, eps = ctx.meta
        N, V = logits2d.shape
        gl = gloss.contiguous().to(torch.float32)
        dl = torch.empty_like(logits2d)
        L.check(L.lib().pka_ce_bwd(L.ptr(logits2d), L.ptr(goal), L.ptr(lse), L.ptr(gl), L.ptr(dl), L.dtype_code(logits2d), N, V,
                                   smoothing, C.c_float(eps), L.stream_ptr()), "ce_bwd")
        return dl, None, None, None


_CE_OUT = []
_CE_DONE = {}          # device index -> uint32[1] CTA-completion counter of the loss kernel (self-resetting)


def cross_entropy_sum(logits2d, goal, smoothing: bool = False, eps: float = 0.1, out3: Optional[torch.Tensor] = None):
    """-> (loss_sum scalar tensor, stats tensor [n_correct, n_words]); PAD (=0) targets ignored (L/train.py:58-90).
    `out3`: caller-owned float[3] the kernel writes {loss, n_correct, n_words} into (the results are views of it)."""
    if out3 is not None:
        _CE_OUT.append(out3)
    try:
        return _CrossEntropyFn.apply(logits2d, goal, bool(smoothing), float(eps))
    finally:
        _CE_OUT.clear()


def split_targets(tgt: torch.Tensor, tgt_pad_mask: torch.Tensor):
    """Teacher-forcing split of L/train.py:163-165 in one launch: -> (tgt[:, :-1], tgt[:, 1:], mask[:, :-1]) as
    contiguous tensors (the three strided views would otherwise each be copied by the op that consumes them)."""
    L.require_cuda(tgt, tgt_pad_mask)
    tgt = tgt.to(torch.int64).contiguous()
    mask = tgt_pad_mask.to(torch.uint8).contiguous()
    B, L1 = tgt.shape
    tgt_in = torch.empty(B, L1 - 1, device=tgt.device, dtype=torch.int64)
    goal = torch.empty(B, L1 - 1, device=tgt.device, dtype=torch.int64)
    mask_in = torch.empty(B, L1 - 1, device=tgt.device, dtype=torch.uint8)
    L.check(L.lib().pka_split_targets(L.ptr(tgt), L.ptr(mask), L.ptr(tgt_in), L.ptr(goal), L.ptr(mask_in), B, L1,
                                      L.stream_ptr()), "split_targets")
    return tgt_in, goal, mask_in


# ------------------------------------------------------------------------------------------------ test helper
def dropout_keep_mask(n: int, drop: Drop, device) -> torch.Tensor:
    """The keep bits (uint8[n]) the kernels derive for this site at the *current* step -- for parity tests only."""
    keep = torch.empty(n, device=device, dtype=torch.uint8)
    L.check(L.lib().pka_dropout_mask(L.ptr(keep), C.c_int64(n), C.byref(drop.c()), L.stream_ptr()), "dropout_mask")
    return keep


def attn_keep_mask(B: int, H: int, Lq: int, Lk: int, drop: Drop, device) -> torch.Tensor:
    """Keep bits [B,H,Lq,Lk] of an attention-probability dropout site: the kernels index element (b,h,i,j) as
    ((b*H+h)*Lq+i)*Lk8 + j with Lk8 = Lk rounded up to a multiple of 8 -- for parity tests only."""
    Lk8 = (Lk + 7) // 8 * 8
    return dropout_keep_mask(B * H * Lq * Lk8, drop, device).view(B, H, Lq, Lk8)[..., :Lk].contiguous()


# ================================================================================================ operand cache
class OperandCache:
    """bf16 GEMM operand copies of a model's fp32 master weights, refreshed by ONE launch per forward pass.

    The tensor-core GEMMs read the weights as bf16 in two layouts (forward: nn.Linear layout; data-gradient: transposed
    per splice segment; per-head projection tensors packed as q|k|v column blocks).  Round 1 rebuilt them inside every
    op (29 small launches per step).  Entries register themselves the first time an op sees a weight while this cache
    is active (one-off relayout with the stand-alone kernels); from then on `refresh()` rewrites all of them with
    pka_relayout_jobs -- the weights only change in the optimiser step -- and, when asked, advances the dropout step
    counter in the same launch.  Entries follow their parameter if its storage moves (FusedAdam arena, .to())."""

    def __init__(self):
        self.entries = {}            # key (storage addresses) -> dict(params, wf, wd, spec, seen)
        self._table = None           # ctypes array, rebuilt when the set of entries changes
        self._n = 0
        self.epoch = 0               # forward passes seen; entries a whole pass did not touch are dropped (moved weights)

    # -- registration (first use)
    def plain(self, weight, N, K, nseg):
        key = ("w", weight.data_ptr(), N, K, nseg)
        ent = self.entries.get(key)
        if ent is not None:
            ent["seen"] = self.epoch
        if ent is None:
            dev = weight.device
            # data-gradient operand [K, N]: its row pitch must be a multiple of 16 bytes for TMA, so an output width
            # that is not a multiple of 8 (V = 53) gets zero columns up to the next one (single-segment layers only)
            Np = (N + 7) // 8 * 8 if nseg == 1 else N
            ent = dict(params=[weight], wf=torch.empty(N, nseg * K, device=dev, dtype=torch.bfloat16),
                       wd=torch.zeros(K, nseg * Np, device=dev, dtype=torch.bfloat16),
                       spec=[(L.RELAYOUT_PLAIN, N, K, nseg, nseg * K, nseg * Np, 0)], seen=self.epoch)
            self.entries[key] = ent
            self._table = None
            self._relayout_now(ent)
        return ent["wf"], ent["wd"]

    def heads(self, ws):
        """Packed operand of the per-head tensors `ws` (each [H, D, dk]): wf [(p,h,j), d], wd [d, (p,h,j)]."""
        key = ("h",) + tuple(w.data_ptr() for w in ws) + tuple(ws[0].shape)
        ent = self.entries.get(key)
        if ent is not None:
            ent["seen"] = self.epoch
        if ent is None:
            H, D, dk = ws[0].shape
            ntot = len(ws) * H * dk
            dev = ws[0].device
            ent = dict(params=list(ws), wf=torch.empty(ntot, D, device=dev, dtype=torch.bfloat16),
                       wd=torch.empty(D, ntot, device=dev, dtype=torch.bfloat16),
                       spec=[(L.RELAYOUT_HEADS, H, D, dk, D, ntot, p_ * H * dk) for p_ in range(len(ws))], seen=self.epoch)
            self.entries[key] = ent
            self._table = None
            self._relayout_now(ent)
        return ent["wf"], ent["wd"]

    def _jobs_of(self, ent):
        jobs = []
        for w, (kind, N, K, nseg, ldf, ldd, n0) in zip(ent["params"], ent["spec"]):
            if not w.is_contiguous() or w.dtype != torch.float32:
                raise RuntimeError("OperandCache: fp32 contiguous master weights expected")
            j = L.RelayoutJob()
            j.src, j.wf, j.wd = w.data_ptr(), ent["wf"].data_ptr(), ent["wd"].data_ptr()
            j.N, j.K, j.nseg, j.kind, j.ldf, j.ldd, j.n0 = N, K, nseg, kind, ldf, ldd, n0
            jobs.append(j)
        return jobs

    def _relayout_now(self, ent):
        jobs = self._jobs_of(ent)
        arr = (L.RelayoutJob * len(jobs))(*jobs)
        L.check(L.lib().pka_relayout_jobs(arr, len(jobs), C.c_void_p(0), L.stream_ptr()), "relayout_jobs")

    # -- once per forward pass
    def refresh(self, step_counter: Optional[torch.Tensor] = None):
        """Rewrite every registered operand copy from its master weight (one launch); `step_counter`: the model's dropout
        step counter, advanced by one in the same launch."""
        stale = [k for k, ent in self.entries.items() if ent["seen"] < self.epoch]
        if stale and self.epoch > 0:                  # weights that moved (or left the model) since the last pass
            for k in stale:
                del self.entries[k]
            self._table = None
        self.epoch += 1
        if self._table is None:
            jobs = [j for ent in self.entries.values() for j in self._jobs_of(ent)]
            self._table = (L.RelayoutJob * max(1, len(jobs)))(*jobs)
            self._n = len(jobs)
        L.check(L.lib().pka_relayout_jobs(self._table, self._n, L.ptr(step_counter), L.stream_ptr()), "relayout_jobs")

    def active(self):
        import contextlib

        @contextlib.contextmanager
        def cm():
            global OPERANDS
            prev, OPERANDS = OPERANDS, self
            try:
                yield self
            finally:
                OPERANDS = prev
        return cm()


OPERANDS: Optional[OperandCache] = None      # set by Transformer.forward around the bf16 path


# ================================================================================================ bf16 tensor-core path
def gemm_tc_rows(A, Bw, Bt, T, N, K, *, nseg=1, lda, ldb, a_seg_col=0, b_seg_col=0, shift=(), bias=None, relu=False,
                 drop: Optional[Drop] = None, out_dtype=torch.bfloat16, addend=None):
    """mode 0 of pka_gemm_tc: C[b,t,:] = epi(sum_seg A[b,t+shift[seg],:] . Bw[:, seg]^T) [+ addend] -> C [Bt,T,N]."""
    L.require_cuda(A, Bw, bias, addend)
    assert A.dtype == torch.bfloat16 and Bw.dtype == torch.bfloat16
    Cout = torch.empty(Bt, T, N, device=A.device, dtype=out_dtype)
    d = L.TcDesc()
    d.A, d.B, d.C, d.Ct = A.data_ptr(), Bw.data_ptr(), Cout.data_ptr(), 0
    d.bias = bias.data_ptr() if bias is not None else 0
    d.mode, d.Bt, d.T, d.Tp, d.M, d.N, d.K, d.nseg = 0, Bt, T, 0, 0, N, K, nseg
    d.lda, d.ldb, d.ldc = lda, ldb, N
    d.a_seg_col, d.b_seg_col = a_seg_col, b_seg_col
    for i, sft in enumerate(shift):
        d.shift[i] = int(sft)
    d.relu, d.c_dtype, d.splits = int(relu), L.dtype_code(Cout), 1
    d.drop = _cdrop(drop)
    if addend is not None:
        assert addend.dtype == torch.bfloat16 and addend.is_contiguous() and addend.numel() == Bt * T * N
        d.addend, d.ldadd = addend.data_ptr(), N
    L.check(L.lib().pka_gemm_tc(C.byref(d), L.stream_ptr()), "gemm_tc")
    return Cout


def gemm_tc_wgrad(dZ, X, Bt, T, M, N, nseg, shift, out=None, accumulate=None, reduce=True, defer=False, heads=None,
                  m_valid=None):
    """mode 2: dW[o, seg*N+i] = sum_{b,t} dZ[b,t,o] * X[b,t+shift[seg],i] -> fp32 [M, nseg*N], straight from the row-major
    activations (MN-major UMMA operands; the frame shift is a TMA row coordinate).  The reduction over all frames is
    split over the utterances to fill the SMs; partial sums are added in a fixed order (deterministic).

    `defer` (backward passes): the fixed-order sum -- and, for problems too small for the CTA-pair kernel, the GEMM
    itself -- is recorded and runs in the batched launches at the end of the backward pass (see flush_deferred).
    `heads` = (grads, P, H, D, dk): the result is the packed gradient [(p,h,j), d] of P per-head projection tensors and
    is summed straight into the reference's per-head layout g_p[h, d, j] (T/SubLayers.py:29-31)."""
    assert dZ.dtype == torch.bfloat16 and X.dtype == torch.bfloat16 and dZ.is_contiguous() and X.is_contiguous()
    pair = M % 256 == 0 and N % 256 == 0
    if pair:          # CTA-pair kernel: 256x256 tiles on two SMs, reduction split in 128-frame units
        tiles = 2 * (M // 256) * (N // 256) * nseg
        units = Bt * ((T + 127) // 128)
        splits = max(1, min(units, 148 // max(1, tiles)))
        splits = -(-units // -(-units // splits))          # no empty splits
    else:
        tiles = ((M + 127) // 128) * ((N + 127) // 128) * nseg
        splits = max(1, min(Bt, 148 // max(1, tiles)))
        if defer and DEFER_ENABLED:
            # grouped launch: the problem does not have to fill the GPU on its own, so a CTA should get at least ~8
            # reduction steps of 64 frames (a split that reduces one 63-token utterance is all prologue and epilogue)
            splits = max(1, min(splits, (Bt * ((T + 63) // 64)) // 8))
        splits = -(-Bt // -(-Bt // splits))                # no empty splits (utterances per split rounds up)
    ws = torch.empty(splits, M, nseg * N, device=dZ.device, dtype=torch.float32)
    d = L.TcDesc()
    d.A, d.B, d.C, d.Ct, d.bias = dZ.data_ptr(), X.data_ptr(), ws.data_ptr(), 0, 0
    d.mode, d.Bt, d.T, d.Tp, d.M, d.N, d.K, d.nseg = 2, Bt, T, 0, M, N, T, nseg
    d.lda, d.ldb, d.ldc = M, N, nseg * N
    for i, sft in enumerate(shift):
        d.shift[i] = int(sft)
    d.relu, d.c_dtype, d.splits = 0, L.PKA_F32, splits
    d.drop = L.NO_DROPOUT
    deferred = defer and _defer_schedule()
    if deferred and not pair:
        with _DEFER.lock:
            _DEFER.wgrads.append(d)
            _DEFER.keep.append((dZ, X, ws))
    elif deferred:                                # a full-wave GEMM nobody waits for: next to the data-gradient chain
        with _SideStream(dZ, X, ws):
            L.check(L.lib().pka_gemm_tc(C.byref(d), L.stream_ptr()), "gemm_tc(wgrad)")
    else:
        L.check(L.lib().pka_gemm_tc(C.byref(d), L.stream_ptr()), "gemm_tc(wgrad)")
    per = M * nseg * N
    if heads is not None:
        grads, P, H, D, dk = heads
        assert M == P * H * dk and N == D and nseg == 1
        if deferred:
            for p_, g in enumerate(grads):
                if g is not None:
                    defer_reduce(ws, g, H * D * dk, splits, per, src_off=p_ * H * dk * D, kind=L.REDUCE_HEADS, D=D, dk=dk)
        elif P <= 3:
            gp = [L.ptr(g) for g in grads] + [C.c_void_p(0)] * (3 - P)
            L.check(L.lib().pka_tc_reduce_heads(L.ptr(ws), gp[0], gp[1], gp[2], splits, P, H, D, dk, L.stream_ptr()),
                    "tc_reduce_heads")
        else:                                      # more than three packed tensors: the job table, launched right away
            jobs = []
            for p_, g in enumerate(grads):
                if g is not None:
                    j = L.ReduceJob()
                    j.src, j.dst = ws.data_ptr() + 4 * p_ * H * dk * D, g.data_ptr()
                    j.n, j.split_stride, j.splits, j.kind, j.accumulate, j.D, j.dk = H * D * dk, per, splits, L.REDUCE_HEADS, 0, D, dk
                    jobs.append(j)
            if jobs:
                arr = (L.ReduceJob * len(jobs))(*jobs)
                L.check(L.lib().pka_reduce_jobs(arr, len(jobs), L.stream_ptr()), "reduce_jobs")
        return grads
    if not reduce:
        return ws, splits                         # the caller sums the split partials itself (fused with its relayout)
    acc = (out is not None) if accumulate is None else bool(accumulate)
    n_out = per if m_valid is None else m_valid * nseg * N          # rows beyond m_valid are zero padding of dZ
    if out is None:
        out = torch.empty(n_out // (nseg * N), nseg * N, device=dZ.device, dtype=torch.float32)
    if deferred:
        defer_reduce(ws, out, n_out, splits, per, accumulate=acc)
    else:
        j = L.ReduceJob()
        j.src, j.dst, j.n, j.split_stride, j.splits, j.kind, j.accumulate = ws.data_ptr(), out.data_ptr(), n_out, per, splits, 0, int(acc)
        L.check(L.lib().pka_reduce_jobs(C.byref(j), 1, L.stream_ptr()), "reduce_jobs")
    return out


def gate_to_bf16(x, Bt, T, N, y=None, scale=1.0):
    """bf16 [Bt,T,N] copy of x (fp32 or bf16); with `y` applies the ReLU+dropout gate (y > 0 ? x*scale : 0)."""
    L.require_cuda(x, y)
    x = x.contiguous()
    out = torch.empty(Bt, T, N, device=x.device, dtype=torch.bfloat16)
    L.check(L.lib().pka_relu_bwd_dual(L.ptr(x), L.dtype_code(x), L.ptr(y), L.ptr(out), C.c_void_p(0), Bt, T, (T + 7) // 8 * 8,
                                      N, C.c_float(scale), int(y is not None), L.stream_ptr()), "relu_bwd_dual")
    return out


def gate_colsum(dy, y, scale, bias_like, Bt, T, N, defer=False):
    """dz = (y > 0 ? dy*scale : 0) as bf16 [Bt,T,N] (y None: plain bf16 copy) AND the bias gradient sum_rows(dz), one pass.
    `defer`: the fixed-order sum of the partial rows runs in the batched launch at the end of the backward pass."""
    L.require_cuda(dy, y)
    dy = dy.contiguous()
    rows = Bt * T
    dz = torch.empty(Bt, T, N, device=dy.device, dtype=torch.bfloat16)
    db = grad_buffer(bias_like)
    chunks = L.lib().pka_colsum_chunks(C.c_int64(rows))
    ws = torch.empty(chunks * N, device=dy.device, dtype=torch.float32)
    deferred = defer and _defer_schedule()
    L.check(L.lib().pka_gate_colsum(L.ptr(dy), L.dtype_code(dy), L.ptr(y), L.ptr(dz), C.c_void_p(0) if deferred else L.ptr(db),
                                    L.ptr(ws), C.c_int64(rows), N, C.c_float(scale), int(y is not None), L.stream_ptr()),
            "gate_colsum")
    if deferred:
        defer_reduce(ws, db, N, L.lib().pka_gate_colsum_parts(L.dtype_code(dy), C.c_int64(rows), N), N)
    return dz, db


def weight_relayout(w2, K, nseg, want_f=True, want_d=True):
    N = w2.shape[0]
    wf = torch.empty(N, nseg * K, device=w2.device, dtype=torch.bfloat16) if want_f else None
    wd = torch.empty(K, nseg * N, device=w2.device, dtype=torch.bfloat16) if want_d else None
    L.check(L.lib().pka_weight_relayout(L.ptr(w2), L.ptr(wf), L.ptr(wd), N, K, nseg, L.stream_ptr()), "weight_relayout")
    return wf, wd


class _LinearTcFn(torch.autograd.Function):
    """bf16 tensor-core version of _LinearFn for [Bt,T,K] activations (tcgen05 forward, data-gradient and
    weight-gradient).  Weights stay fp32 masters; their bf16 operand copies are made here, once per call."""

    @staticmethod
    def forward(ctx, x, weight, bias, splice, relu, drop, out_fp32, link=None, single_consumer=False):
        L.require_cuda(x, weight, bias)
        assert x.dtype == torch.bfloat16 and x.dim() == 3 and x.is_contiguous()
        Bt, T, kin = x.shape
        n_ctx = len(splice) if splice else 1
        w2 = weight.reshape(weight.shape[0], -1)
        N = w2.shape[0]
        assert w2.shape[1] == n_ctx * kin
        needs_dx = ctx.needs_input_grad[0]
        if OPERANDS is not None and weight.is_contiguous():
            wf, wd = OPERANDS.plain(weight, N, kin, n_ctx)         # refreshed once per forward pass, not per op
        elif N % 8 != 0 and n_ctx == 1 and weight.is_contiguous():
            wf, wd = OperandCache().plain(weight, N, kin, 1)       # one-off entry: zero-padded data-gradient operand
        else:
            wf, wd = weight_relayout(w2.detach(), kin, n_ctx, True, needs_dx)
        if link is not None:
            link.armed = bool(needs_dx) and not splice
        ctx.link = link if (link is not None and link.armed) else None
        y = gemm_tc_rows(x, wf, Bt, T, N, kin, nseg=n_ctx, lda=kin, ldb=n_ctx * kin, b_seg_col=kin, shift=splice or (),
                         bias=bias, relu=relu, drop=drop, out_dtype=torch.float32 if out_fp32 else torch.bfloat16)
        if relu and GATE_TAP is not None:
            GATE_TAP.append(y)
        # dropout of the INPUT (applied by the op that produced x) folded into this layer's data-gradient epilogue
        rec = _IN_DROP.get(x.data_ptr())
        ctx.in_drop = None
        if rec is not None and needs_dx:
            rec.fused = True
            ctx.in_drop = rec.drop
        # ... and this layer's own output dropout offered to ITS single consumer (no ReLU: the gate kernel handles that case)
        ctx.rec = None
        if single_consumer and not relu and drop is not None and drop.on and y.dtype == torch.bfloat16:
            ctx.rec = _IN_DROP[y.data_ptr()] = _DropRec(drop)
        ctx.save_for_backward(x, wd, y if relu else None, w2, bias)
        ctx.meta = (Bt, T, kin, N, n_ctx, tuple(splice) if splice else (0,), relu, drop, bias is not None, weight.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wd, y, w2, bias = ctx.saved_tensors
        Bt, T, kin, N, n_ctx, splice, relu, drop, has_bias, wshape = ctx.meta
        dy = dy.contiguous()
        use_drop = drop is not None and drop.on
        dx = dw = db = None
        want_db = has_bias and ctx.needs_input_grad[2]
        if relu:
            assert y.dtype == torch.bfloat16
            scale = (1.0 / (1.0 - drop.p)) if use_drop else 1.0
            if want_db and N % 4 == 0:            # ReLU/dropout gate and bias gradient in one pass over dY
                dz, db = gate_colsum(dy, y, scale, bias, Bt, T, N, defer=True)
                want_db = False
            else:
                dz = gate_to_bf16(dy, Bt, T, N, y=y, scale=scale)
        else:
            if use_drop and not (ctx.rec is not None and ctx.rec.fused):
                tmp = torch.empty_like(dy)
                L.check(L.lib().pka_dropout_bwd(L.ptr(dy), L.ptr(tmp), L.dtype_code(dy), C.c_int64(dy.numel()),
                                                _byref_drop(drop), L.stream_ptr()), "dropout_bwd")
                dy = tmp
            if dy.dtype == torch.bfloat16 or N % 8 != 0:
                dz = dy                           # (N % 8 != 0: converted and zero-padded below)
            elif want_db and N % 4 == 0:          # fp32 -> bf16 copy and bias gradient in one pass
                dz, db = gate_colsum(dy, None, 1.0, bias, Bt, T, N, defer=True)
                want_db = False
            else:
                dz = gate_to_bf16(dy, Bt, T, N)
        addend = ctx.link.take() if ctx.link is not None else None       # residual-branch gradient of the sub-layer input
        Np = N
        if N % 8 != 0:                            # V = 53: zero-pad the gradient to a TMA-legal width (wd is padded alike)
            assert n_ctx == 1 and wd.shape[1] % 8 == 0 and not has_bias
            Np = wd.shape[1]
            dzp = torch.empty(Bt, T, Np, device=dy.device, dtype=torch.bfloat16)
            L.check(L.lib().pka_pad_cast(L.ptr(dy), L.dtype_code(dy), L.ptr(dzp), C.c_int64(Bt * T), N, Np, L.stream_ptr()),
                    "pad_cast")
            dz = dzp
        if ctx.needs_input_grad[0]:
            dx = gemm_tc_rows(dz, wd, Bt, T, kin, Np, nseg=n_ctx, lda=Np, ldb=n_ctx * Np, b_seg_col=Np,
                              shift=[-c for c in splice], addend=addend, drop=ctx.in_drop)
        if ctx.needs_input_grad[1]:
            dw = gemm_tc_wgrad(dz, x, Bt, T, Np, kin, n_ctx, splice, out=grad_buffer(w2, (N, n_ctx * kin)),
                               accumulate=False, defer=True, m_valid=N).view(wshape)
        if want_db:
            hint = _COLSUM_HINTS.pop(dz.data_ptr(), None) if dz is dy else None
            if hint is not None and hint[2] == Bt * T and hint[3] == N and _defer_schedule():
                db = grad_buffer(bias)            # column sums of dz were left behind by the LayerNorm backward above us
                defer_reduce(hint[0], db, N, hint[1], 3 * N, src_off=2 * N)
            else:
                db = colsum(dz.view(Bt * T, N), out=grad_buffer(bias), accumulate=False, defer=True)
        return dx, dw, db, None, None, None, None, None, None


GATE_TAP = None          # parity tests set this to a list: every [ReLU] tensor-core layer appends its output (call order)


def linear_tc(x, weight, bias=None, splice=None, relu=False, drop=None, out_fp32=False, link: Optional[ResidualLink] = None,
              single_consumer: bool = False):
    return _LinearTcFn.apply(x, weight, bias, list(splice) if splice else None, relu, drop, out_fp32, link, single_consumer)


def transpose_to_bf16(weight_kn):
    """[K, N] fp32 -> [N, K] bf16 (nn.Linear layout of a matrix stored input-major, e.g. the LDA transform)."""
    K, N = weight_kn.shape
    wt = torch.empty(N, K, device=weight_kn.device, dtype=torch.bfloat16)
    L.check(L.lib().pka_transpose(L.ptr(weight_kn.detach().contiguous()), L.PKA_F32, L.ptr(wt), L.PKA_BF16, K, N, L.stream_ptr()),
            "transpose")
    return wt


def affine_tc(x, weight_kn, bias=None, wt=None):
    """Frozen LDA affine on the tensor cores: x bf16 [Bt,T,K] @ W[K,N] + b -> bf16.  No autograd.  `wt` = cached
    transpose_to_bf16(weight_kn) (the matrix is frozen, so callers keep it across steps)."""
    Bt, T, K = x.shape
    N = weight_kn.shape[1]
    if wt is None:
        wt = transpose_to_bf16(weight_kn)
    return gemm_tc_rows(x.detach(), wt, Bt, T, N, K, lda=K, ldb=K, bias=bias.detach() if bias is not None else None)
