"""GPU diagnostic + micro-benchmark for the tcgen05 GEMM path at the bench shape (Bt=32, T=499, 256 -> 256, 3 contexts)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_kaldi_asr_b200 import ops
torch.manual_seed(0)
dev = "cuda"
Bt, T, K, N, ctx = 32, 499, 256, 256, [-3, 0, 3]
x = torch.randn(Bt, T, K, device=dev).bfloat16()
dz = torch.randn(Bt, T, N, device=dev).bfloat16()
w = torch.randn(N, 3 * K, device=dev) / (3 * K) ** 0.5
b = torch.randn(N, device=dev) * 0.1
wf, wd = ops.weight_relayout(w, K, 3)
xs = [x.clone() for _ in range(10)]
dzs = [dz.clone() for _ in range(10)]
it = [0]
def fwd():
    it[0] += 1
    return ops.gemm_tc_rows(xs[it[0] % 10], wf, Bt, T, N, K, nseg=3, lda=K, ldb=3 * K, b_seg_col=K, shift=ctx, bias=b, relu=True)
def dgrad():
    it[0] += 1
    return ops.gemm_tc_rows(dzs[it[0] % 10], wd, Bt, T, K, N, nseg=3, lda=N, ldb=3 * N, b_seg_col=N, shift=[-c for c in ctx])
def wgrad():
    it[0] += 1
    return ops.gemm_tc_wgrad(dzs[it[0] % 10], xs[it[0] % 10], Bt, T, N, K, 3, ctx)
flops = 2.0 * Bt * T * 768 * 256
for nm, fn in (("fwd", fwd), ("dgrad", dgrad), ("wgrad", wgrad)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): g.replay()
    e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) / 100 * 1e3
    print("%s: %.1f us per call (CUDA-graph replay, device time), %.1f TFLOP/s" % (nm, us, flops / us / 1e6), flush=True)
