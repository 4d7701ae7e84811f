"""GPU diagnostic: pka_gemm_f32 layouts x tile configs vs torch matmul."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_kaldi_asr_b200 import ops
torch.manual_seed(0)
dev = "cuda"
def err(a, b): return float((a - b).abs().max() / b.abs().max())
for tile in ("small", "big"):
    os.environ["PKA_GEMM_TILE"] = tile
    for (M, N, K) in ((1290, 256, 256), (13760, 256, 256), (300, 200, 96)):
        A = torch.randn(M, K, device=dev); Bt = torch.randn(N, K, device=dev); Bn = torch.randn(K, N, device=dev); At = torch.randn(K, M, device=dev)
        C = torch.empty(M, N, device=dev)
        ops.gemm(A, Bt, C, M, N, K, lda=K, ldb=K, ldc=N, transA=False, transB=True); e1 = err(C, A @ Bt.t())
        ops.gemm(A, Bn, C, M, N, K, lda=K, ldb=N, ldc=N, transA=False, transB=False); e2 = err(C, A @ Bn)
        ops.gemm(At, Bn, C, M, N, K, lda=M, ldb=N, ldc=N, transA=True, transB=False); e3 = err(C, At.t() @ Bn)
        ops.gemm(At, Bt, C, M, N, K, lda=M, ldb=K, ldc=N, transA=True, transB=True); e4 = err(C, At.t() @ Bt.t())
        print(tile, (M, N, K), "NT %.1e NN %.1e TN %.1e TT %.1e" % (e1, e2, e3, e4))
    # spliced dgrad shape
    T, Bsz, D = 430, 32, 256
    M = T * Bsz
    dz = torch.randn(M, D, device=dev); W = torch.randn(D, 3 * D, device=dev)
    dx = torch.empty(M, D, device=dev)
    ops.gemm(dz, W, dx, M, D, D, nseg=3, lda=D, ldb=3 * D, ldc=D, transB=False, b_seg_off=D, shiftA=[3, 0, -3], T=T)
    ref = torch.zeros(Bsz, T, D, device=dev)
    dz3 = dz.view(Bsz, T, D)
    for s, c in enumerate([-3, 0, 3]):
        contrib = dz3 @ W[:, s * D:(s + 1) * D]          # contribution of output rows t to input rows t+c
        lo, hi = max(0, -c), min(T, T - c)
        ref[:, lo + c:hi + c] += contrib[:, lo:hi]
    print(tile, "spliced dgrad", "%.1e" % err(dx.view(Bsz, T, D), ref))
    for sidx, sh in enumerate([3, 0, -3]):
        ops.gemm(dz, W, dx, M, D, D, nseg=1, lda=D, ldb=3 * D, ldc=D, transB=False, shiftA=[sh], T=T, b_ptr_off=sidx * D)
        c = -sh
        r = torch.zeros(Bsz, T, D, device=dev)
        contrib = dz3 @ W[:, sidx * D:(sidx + 1) * D]
        lo, hi = max(0, -c), min(T, T - c)
        r[:, lo + c:hi + c] = contrib[:, lo:hi]
        print(tile, "  single seg shift", sh, "%.1e" % err(dx.view(Bsz, T, D), r))
