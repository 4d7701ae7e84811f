"""Times the tensor-core attention kernels alone (CUDA-graph timed, like bench.py's probes) at the config-5 encoder shape
and at the TIMIT decoder shapes.  usage: python tools/bench_attn.py"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import time_kernel, _band_pairs, peaks
from pytorch_kaldi_asr_b200 import ops

def run(name, B, H, Lq, Lk, band, p_drop, d_model):
    HD = H * 64
    cross = Lq != Lk
    mask = torch.ones(B, Lk, device="cuda", dtype=torch.uint8)
    step = torch.zeros(1, device="cuda", dtype=torch.int64)
    drop = ops.Drop(p_drop, 3, 7, step) if p_drop > 0 else None
    if cross:
        q = [(torch.randn(B, Lq, HD, device="cuda") * 0.5).bfloat16().requires_grad_(True) for _ in range(4)]
        kv = [(torch.randn(B, Lk, 2 * HD, device="cuda") * 0.5).bfloat16().requires_grad_(True) for _ in range(4)]
    else:
        q = [(torch.randn(B, Lq, 3 * HD, device="cuda") * 0.5).bfloat16().requires_grad_(True) for _ in range(4)]
        kv = [None] * 4
    gy = (torch.randn(B, Lq, HD, device="cuda")).bfloat16()
    i = [0]
    def fwd():
        i[0] += 1
        with torch.no_grad():
            ops.attention_tc(q[i[0] % 4], kv[i[0] % 4], mask, H, 64, band, 1.0 / math.sqrt(d_model), drop)
    def both():                                   # forward + backward inside one capture (autograd replays the backward
        i[0] += 1                                 # on the forward's stream, which must be the capturing one)
        out, _ = ops.attention_tc(q[i[0] % 4], kv[i[0] % 4], mask, H, 64, band, 1.0 / math.sqrt(d_model), drop)
        out.backward(gy)
    tf = time_kernel(fwd, iters=8, replays=3)
    tb = time_kernel(both, iters=4, replays=3) - tf
    pairs = B * H * (_band_pairs(Lq, Lk, band) if not cross else Lq * Lk)
    fl = pairs * 4.0 * 64
    print("%-34s fwd %8.2f us %7.1f TFLOP/s (%.1f%% of burst) | bwd(dQ + dK/dV) %8.2f us %7.1f TFLOP/s"
          % (name, tf * 1e6, fl / tf / 1e12, 100 * fl / tf / 1e12 / peaks()["bf16_tflops"], tb * 1e6, 2.5 * fl / tb / 1e12), flush=True)

run("cfg5 enc full  B4 H8 T1600", 4, 8, 1600, 1600, None, 0.0, 512)
run("cfg5 enc full  B4 H8 T1600 p=.1", 4, 8, 1600, 1600, None, 0.1, 512)
run("cfg5 enc band(-100,0)", 4, 8, 1600, 1600, (-100, 0), 0.1, 512)
run("timit dec self  B32 H2 L63", 32, 2, 63, 63, (-10, 0), 0.35, 128)
run("timit dec cross B32 H2 63x499", 32, 2, 63, 499, None, 0.35, 128)
