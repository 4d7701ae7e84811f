"""Per-CTA timeline of the A-stationary rows GEMM (PKA_TC_DBG=16 stamps), TDNN shape."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pytorch_kaldi_asr_b200 import ops, _lib
B, T = 32, 499
x = torch.randn(B, T, 256, device="cuda").bfloat16()
w = torch.randn(256, 768, device="cuda") * 0.03
wf, wd = ops.weight_relayout(w, 256, 3)
bias = torch.randn(256, device="cuda")
os.environ["PKA_TC_DBG"] = sys.argv[1] if len(sys.argv) > 1 else "16"
if len(sys.argv) > 2: os.environ["PKA_TC_CLUSTER"] = sys.argv[2]
for _ in range(3):
    ops.gemm_tc_rows(x, wf, B, T, 256, 256, nseg=3, lda=256, ldb=768, b_seg_col=256, shift=[-3, 0, 3], bias=bias, relu=True)
torch.cuda.synchronize()
buf = np.zeros(256 * 16, dtype=np.uint64)
_lib.lib().pka_debug_rows2_stamps(buf.ctypes.data_as(C.c_void_p), buf.nbytes)
ts = buf.reshape(256, 16)[:128].astype(np.int64)
t0 = ts[:, 0].min()
names = ["start", "setup done", "-", "-", "producer done", "A blk0,1 landed", "B st0 landed", "mma nt0 issued", "mma nt1 issued", "-",
         "acc0 ready", "acc1 ready", "last acc ready", "-", "epi done", "exit"]
rel = ts - t0
for k, n in enumerate(names):
    if n != "-":
        print("%-14s min %7d  median %7d  max %7d ns" % (n, rel[:, k].min(), np.median(rel[:, k]), rel[:, k].max()))
print("per-CTA duration median %d ns" % np.median(ts[:, 15] - ts[:, 0]))
