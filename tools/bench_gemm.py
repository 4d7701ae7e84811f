"""Micro-benchmark of the TDNN GEMM family at the TIMIT step's shape under diagnostic env switches (graph-timed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pytorch_kaldi_asr_b200 import ops
B, T = 32, 499
ctx = [-3, 0, 3]
xs = [torch.randn(B, T, 256, device="cuda").bfloat16() for _ in range(24)]
w = torch.randn(256, 768, device="cuda") * 0.03
wf, wd = ops.weight_relayout(w, 256, 3)
bias = torch.randn(256, device="cuda")
i = [0]
def fwd():
    i[0] += 1
    ops.gemm_tc_rows(xs[i[0] % 24], wf, B, T, 256, 256, nseg=3, lda=256, ldb=768, b_seg_col=256, shift=ctx, bias=bias, relu=True)
step = torch.zeros(1, dtype=torch.int64, device="cuda")
drop = ops.Drop(0.35, 3, 1234, step)
def fwd_drop():
    i[0] += 1
    ops.gemm_tc_rows(xs[i[0] % 24], wf, B, T, 256, 256, nseg=3, lda=256, ldb=768, b_seg_col=256, shift=ctx, bias=bias, relu=True, drop=drop)
def plain():      # same FLOPs without the splice: K = 768 contiguous
    i[0] += 1
    ops.gemm_tc_rows(xs3[i[0] % 8], wf, B, T, 256, 768, lda=768, ldb=768)
def wgrad():
    i[0] += 1
    ops.gemm_tc_wgrad(xs[i[0] % 24], xs[(i[0] + 7) % 24], B, T, 256, 256, 3, ctx)
flops = 2.0 * B * T * 768 * 256
variants = [a.split(",") for a in sys.argv[1:]] or [["PKA_TC_ROWS2=0"], ["PKA_TC_ROWS2=1", "PKA_TC_CLUSTER=1"], ["PKA_TC_ROWS2=1", "PKA_TC_CLUSTER=2"],
            ["PKA_TC_ROWS2=1", "PKA_TC_CLUSTER=4"], ["PKA_TC_CLUSTER=4", "PKA_TC_DBG=1"], ["PKA_TC_CLUSTER=4", "PKA_TC_DBG=8"],
            ["PKA_TC_CLUSTER=4", "PKA_TC_DBG=9"], ["PKA_TC_CLUSTER=1", "PKA_TC_DBG=9"], ["PKA_TC_CLUSTER=4", "PKA_TC_DBG=4"]]
for v in variants:
    for kv in v:
        k, val = kv.split("=")
        os.environ[k] = val
    sec = bench.time_kernel(fwd)
    sec2 = bench.time_kernel(fwd_drop)
    print("%-45s fwd %7.2f us  %7.1f TFLOP/s   with dropout %7.2f us" % (" ".join(v), sec * 1e6, flops / sec / 1e12, sec2 * 1e6), flush=True)
    for kv in v:
        os.environ.pop(kv.split("=")[0])
sec = bench.time_kernel(wgrad)
print("wgrad %7.2f us  %7.1f TFLOP/s" % (sec * 1e6, flops / sec / 1e12))
