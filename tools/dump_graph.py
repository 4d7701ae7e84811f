"""Dump the CUDA graph of the training step to a .dot file and count programmatic (PDL) edges."""
import os, sys, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_kaldi_asr_b200 as pk
from pytorch_kaldi_asr_b200 import ops
from pytorch_kaldi_asr_b200.utils import synthetic
pk.set_compute_mode("bf16")
os.makedirs("gpurun_out", exist_ok=True)
x = torch.randn(4, 499, 256, device="cuda").bfloat16()
w = torch.randn(256, 768, device="cuda") * 0.03
wf, wd = ops.weight_relayout(w, 256, 3)
for _ in range(2):
    ops.gemm_tc_rows(x, wf, 4, 499, 256, 256, nseg=3, lda=256, ldb=768, b_seg_col=256, shift=[-3, 0, 3])
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
g.enable_debug_mode()
with torch.cuda.graph(g):
    for _ in range(3):
        y = ops.gemm_tc_rows(x, wf, 4, 499, 256, 256, nseg=3, lda=256, ldb=768, b_seg_col=256, shift=[-3, 0, 3])
        z = ops.add_pos_dropout(y, None, ops.Drop(0.1, 1, 2, torch.zeros(1, dtype=torch.int64, device="cuda")))
g.debug_dump("gpurun_out/graph_small.dot")
txt = open("gpurun_out/graph_small.dot").read()
print("nodes", txt.count("label="), "edges", txt.count("->"), "programmatic mentions", len(re.findall(r"(?i)programmatic|port", txt)))
print(txt[:3000])
