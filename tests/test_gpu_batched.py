"""Batched launches of the training step (`-m gpu`): the one-launch reduction table, the one-launch operand refresh, the
grouped weight-gradient GEMM, the residual-gradient epilogue, the teacher-forcing split -- each against the stand-alone
kernels / plain torch arithmetic it replaces -- and the whole backward pass with and without deferral."""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale).to(DEV)


def test_reduce_jobs_plain_and_heads_match_fixed_order_sums():
    from pytorch_kaldi_asr_b200 import _lib as L
    jobs, want, keep = [], [], []
    for i, (n, splits) in enumerate([(196608, 24), (256, 250), (128, 64), (130, 3), (7, 1), (512, 1184), (4096, 17)]):
        src = rnd(splits, n, seed=i)
        dst = torch.full((n,), 7.0, device=DEV)
        acc = i % 2 == 1
        j = L.ReduceJob()
        j.src, j.dst, j.n, j.split_stride, j.splits, j.kind, j.accumulate = src.data_ptr(), dst.data_ptr(), n, n, splits, 0, int(acc)
        jobs.append(j)
        want.append(src.double().sum(0) + (7.0 if acc else 0.0))
        keep.append((src, dst))
    H, D, dk, P, splits = 2, 128, 64, 3, 5
    packed = rnd(splits, P * H * dk, D, seed=50)
    heads = [torch.zeros(H, D, dk, device=DEV) for _ in range(P)]
    for p in range(P):
        j = L.ReduceJob()
        j.src, j.dst = packed.data_ptr() + 4 * p * H * dk * D, heads[p].data_ptr()
        j.n, j.split_stride, j.splits, j.kind, j.D, j.dk = H * D * dk, P * H * dk * D, splits, 1, D, dk
        jobs.append(j)
    arr = (L.ReduceJob * len(jobs))(*jobs)
    L.check(L.lib().pka_reduce_jobs(arr, len(jobs), L.stream_ptr()), "reduce_jobs")
    torch.cuda.synchronize()
    for (src, dst), w in zip(keep, want):
        assert float((dst.double() - w).abs().max()) <= 1e-5 * max(1.0, float(w.abs().max()))
    summed = packed.double().sum(0).view(P, H, dk, D)
    for p in range(P):
        assert float((heads[p].double() - summed[p].permute(0, 2, 1)).abs().max()) <= 1e-5
    first = [dst.clone() for _, dst in keep]
    for _, dst in keep:
        dst.fill_(7.0)
    L.check(L.lib().pka_reduce_jobs(arr, len(jobs), L.stream_ptr()), "reduce_jobs")
    torch.cuda.synchronize()
    assert all(torch.equal(a, dst) for a, (_, dst) in zip(first, keep)), "summation order must not depend on the launch"


def test_operand_cache_matches_the_stand_alone_relayout_kernels_and_ticks_the_counter():
    from pytorch_kaldi_asr_b200 import _lib as L, ops
    cache = ops.OperandCache()
    w = rnd(256, 768, seed=1, scale=0.05)
    ws = [rnd(2, 128, 64, seed=10 + p, scale=0.1) for p in range(3)]
    wf, wd = cache.plain(w, 256, 256, 3)
    hf, hd = cache.heads(ws)
    rf, rd = ops.weight_relayout(w, 256, 3)
    assert torch.equal(wf, rf) and torch.equal(wd, rd)
    want_f = torch.empty(3 * 2 * 64, 128, device=DEV, dtype=torch.bfloat16)
    want_d = torch.empty(128, 3 * 2 * 64, device=DEV, dtype=torch.bfloat16)
    L.check(L.lib().pka_head_weight_relayout(L.ptr(ws[0]), L.ptr(ws[1]), L.ptr(ws[2]), 3, 2, 128, 64, L.ptr(want_f), L.ptr(want_d),
                                             L.stream_ptr()), "head_weight_relayout")
    assert torch.equal(hf, want_f) and torch.equal(hd, want_d)
    counter = torch.tensor([41], device=DEV, dtype=torch.int64)
    w.mul_(2.0)
    ws[1].add_(1.0)
    cache.refresh(counter)
    cache.plain(w, 256, 256, 3), cache.heads(ws)                      # "used in this pass"
    assert int(counter) == 42
    rf, rd = ops.weight_relayout(w, 256, 3)
    assert torch.equal(wf, rf) and torch.equal(wd, rd)
    assert torch.equal(hf[128:256], ws[1].permute(0, 2, 1).reshape(128, 128).bfloat16())
    # a weight whose storage moved gets a fresh entry; the stale one is dropped after a pass that did not touch it
    w2 = w.clone()
    cache.refresh(None)
    f2, _ = cache.plain(w2, 256, 256, 3)
    cache.heads(ws)
    cache.refresh(None)
    assert len(cache.entries) == 2 and torch.equal(f2, w2.bfloat16())


def test_grouped_weight_gradients_equal_single_launches():
    from pytorch_kaldi_asr_b200 import _lib as L, ops
    shapes = [(32, 63, 384, 128), (32, 63, 128, 128), (8, 499, 256, 128), (3, 130, 128, 256), (32, 63, 128, 128)]
    descs, outs, singles = [], [], []
    for i, (Bt, T, M, N) in enumerate(shapes):
        dz, x = rnd(Bt, T, M, seed=i).bfloat16(), rnd(Bt, T, N, seed=100 + i).bfloat16()
        ws, splits = ops.gemm_tc_wgrad(dz, x, Bt, T, M, N, 1, (0,), reduce=False)
        singles.append(ws.sum(0))
        ws2 = torch.zeros_like(ws)
        d = L.TcDesc()
        d.A, d.B, d.C = dz.data_ptr(), x.data_ptr(), ws2.data_ptr()
        d.mode, d.Bt, d.T, d.M, d.N, d.K, d.nseg, d.lda, d.ldb, d.ldc = 2, Bt, T, M, N, T, 1, M, N, N
        d.c_dtype, d.splits = L.PKA_F32, splits
        d.drop = L.NO_DROPOUT
        descs.append(d)
        outs.append((ws, ws2, dz, x))
    arr = (L.TcDesc * len(descs))(*descs)
    L.check(L.lib().pka_gemm_tc_wgrad_group(arr, len(descs), L.stream_ptr()), "gemm_tc_wgrad_group")
    torch.cuda.synchronize()
    for (ws, ws2, dz, x), (Bt, T, M, N) in zip(outs, shapes):
        assert torch.equal(ws, ws2), "grouped launch differs from the single launch for %r" % ((Bt, T, M, N),)
        ref = torch.einsum("bto,bti->oi", dz.float(), x.float())
        assert float((ws2.sum(0) - ref).abs().max() / ref.abs().max()) <= 1e-2


@pytest.mark.parametrize("Bt,T,K,N", [(32, 63, 384, 128), (4, 499, 128, 128), (3, 300, 512, 512), (2, 77, 128, 56)])
def test_data_gradient_gemm_adds_the_residual_branch_in_its_epilogue(Bt, T, K, N):
    from pytorch_kaldi_asr_b200 import ops
    dz = rnd(Bt, T, K, seed=1).bfloat16()
    wd = rnd(N, K, seed=2, scale=1.0 / math.sqrt(K)).bfloat16()
    add = rnd(Bt, T, N, seed=3).bfloat16()
    plain = ops.gemm_tc_rows(dz, wd, Bt, T, N, K, lda=K, ldb=K, out_dtype=torch.float32)
    fused = ops.gemm_tc_rows(dz, wd, Bt, T, N, K, lda=K, ldb=K, addend=add)
    want = (plain + add.float()).bfloat16()
    assert torch.equal(fused, want)


def test_split_targets_is_the_teacher_forcing_split():
    from pytorch_kaldi_asr_b200 import ops
    tgt = torch.randint(0, 53, (7, 23), device=DEV)
    mask = (torch.rand(7, 23, device=DEV) > 0.3).to(torch.uint8)
    a, b, c = ops.split_targets(tgt, mask)
    assert torch.equal(a, tgt[:, :-1]) and torch.equal(b, tgt[:, 1:]) and torch.equal(c, mask[:, :-1])
    assert a.is_contiguous() and b.is_contiguous() and c.is_contiguous()


def _grads(defer, dropout):
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200 import ops
    from pytorch_kaldi_asr_b200.utils import synthetic
    from oracle import acoustic_model as am
    cfg = am.example_config(en_dropout=dropout, de_dropout=dropout)
    lda = synthetic.lda_matrix()
    sd = am.init_state_dict(cfg, lda, seed=0)
    batch = synthetic.batches(1, 5, seed=31)[0]
    model = pk.Transformer(lda_mat=lda, seed=3, **{k: v for k, v in cfg.items() if k != "encoder_type"})
    model.load_state_dict(sd)
    model = model.to(DEV).train()
    opt = pk.FusedAdam(model.parameters())
    opt.zero_grad()
    old = ops.DEFER_ENABLED
    ops.DEFER_ENABLED = defer
    pk.set_compute_mode("bf16")
    try:
        n0 = pk._lib.launch_count()
        src, smask, tgt, tmask = pk.train._to_device(batch, DEV)
        tgt_in, goal, tmask_in = ops.split_targets(tgt, tmask)
        pred = model(src, smask, tgt_in, tmask_in)
        loss, _ = ops.cross_entropy_sum(pred.view(-1, pred.size(-1)), goal.view(-1), False)
        loss.backward()
        torch.cuda.synchronize()
        launches = pk._lib.launch_count() - n0
    finally:
        pk.set_compute_mode("fp32")
        ops.DEFER_ENABLED = old
    return float(loss.detach()), opt.flat_grad.clone(), launches


@pytest.mark.parametrize("dropout", [0.0, 0.35])
def test_backward_with_deferred_reductions_equals_one_launch_per_reduction(dropout):
    """Same step (same Philox masks) with the finishing work batched at the end of backward vs. launched op by op: equal
    losses, gradients equal to fp32 summation-order noise, far fewer launches; and the batched pass is bit-reproducible."""
    l0, g0, n0 = _grads(False, dropout)
    l1, g1, n1 = _grads(True, dropout)
    l2, g2, _ = _grads(True, dropout)
    assert l0 == l1 == l2
    assert torch.equal(g1, g2)
    rel = float((g0 - g1).abs().max() / g0.abs().max())
    assert rel <= 2e-6, rel
    assert n1 <= n0 - 60, "deferral should remove >= 60 launches per step (%d -> %d)" % (n0, n1)
