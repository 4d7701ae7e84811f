"""Runs every hot kernel of the path once at a representative (large enough) shape so that one `ncu --set full` pass can
collect the counters the roofline claims rest on: tensor-pipe utilisation for the GEMM / attention kernels, achieved
DRAM bandwidth for the norm, front-end, optimiser and top-k kernels.  usage: python tools/kernel_evidence.py [what ...]
what in {gemm, attn, ln, frontend, adam, decode}; default = all."""
import math, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pytorch_kaldi_asr_b200 as pk
from pytorch_kaldi_asr_b200 import ops
from pytorch_kaldi_asr_b200.utils import synthetic

what = sys.argv[1:] or ["gemm", "attn", "ln", "frontend", "adam", "decode"]
dev = "cuda"
REP = 3

if "gemm" in what:          # TDNN layer at the TIMIT step's shape: [32*499, 3x256] x [768, 256]
    B, T = 32, 499
    x = torch.randn(B, T, 256, device=dev).bfloat16()
    dz = torch.randn(B, T, 256, device=dev).bfloat16()
    w = torch.randn(256, 768, device=dev) * 0.03
    wf, wd = ops.weight_relayout(w, 256, 3)
    bias = torch.randn(256, device=dev)
    for _ in range(REP):
        ops.gemm_tc_rows(x, wf, B, T, 256, 256, nseg=3, lda=256, ldb=768, b_seg_col=256, shift=[-3, 0, 3], bias=bias, relu=True)
        ops.gemm_tc_rows(dz, wd, B, T, 256, 256, nseg=3, lda=256, ldb=768, b_seg_col=256, shift=[3, 0, -3])
        ops.gemm_tc_wgrad(dz, x, B, T, 256, 256, 3, [-3, 0, 3])

if "attn" in what:          # config-5 encoder self-attention: B=4, H=8, T=1599, d=64; band (-100,0) and full
    B, H, T, D = 4, 8, 1599, 64
    qkv = torch.randn(B, T, 3 * H * D, device=dev).bfloat16().requires_grad_(True)
    km = torch.ones(B, T, dtype=torch.uint8, device=dev)
    gy = torch.randn(B, T, H * D, device=dev).bfloat16()
    for band in ((-100, 0), None):
        for _ in range(REP):
            out, _ = ops.attention_tc(qkv, None, km, H, D, band, 1.0 / math.sqrt(512.0), None, False)
            out.backward(gy)
            qkv.grad = None

if "ln" in what:            # residual add + LayerNormalization over 32*499*14 rows of 256 (bf16 and fp32), forward + backward
    rows = 32 * 499 * 14
    for dt in (torch.bfloat16, torch.float32):
        x = torch.randn(1, rows, 256, device=dev).to(dt).requires_grad_(True)
        r = torch.randn(1, rows, 256, device=dev).to(dt)
        a = torch.ones(256, device=dev, requires_grad=True)
        b = torch.zeros(256, device=dev, requires_grad=True)
        gy = torch.randn(1, rows, 256, device=dev).to(dt)
        for _ in range(REP):
            y = ops.add_layer_norm(x, r, a, b)
            y.backward(gy)
            x.grad = None

if "frontend" in what:      # CMVN + splice(+-2) on 2048 padded utterances of 499 frames x 40 (163 MB in, 409 MB out)
    feats = torch.randn(2048, 499, 40, device=dev)
    lengths = torch.full((2048,), 450, dtype=torch.int32, device=dev)
    for _ in range(REP):
        ops.frontend(feats, lengths, 1, [-2, -1, 0, 1, 2], 1, out_dtype=torch.bfloat16)
        ops.frontend(feats, None, 1, [-2, -1, 0, 1, 2], 0, out_dtype=torch.bfloat16)

if "adam" in what:          # fused Adam over a 35 M parameter arena (config 5's size)
    p = torch.nn.Parameter(torch.randn(35_000_000, device=dev))
    opt = pk.FusedAdam([p])
    p.grad = torch.randn_like(p)
    for _ in range(REP):
        opt.step()

if "decode" in what:        # beam-10 decoding of 250 utterances: tree attention over the KV cache + warp top-k lattice update
    from pytorch_kaldi_asr_b200.decode import translate_batch
    from pytorch_kaldi_asr_b200.utils.instances_handler import pad_to_longest
    pk.set_compute_mode("fp32")
    cfg = dict(n_src_dim=40, n_tgt_vocab=53, encoder_max_len=500, decoder_max_len=100, src_fold=1, encoder_sub_sequence=(-100, 0),
               decoder_sub_sequence=(-10, 0), en_layers=3, de_layers=3, n_head=2, en_d_model=256, de_d_model=128, d_k=64, d_v=64,
               en_dropout=0.35, de_dropout=0.35, tdnn_contexts=[[-1, 0, 1], [-1, 0, 1], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3]])
    torch.manual_seed(0)
    model = pk.Transformer(lda_mat=synthetic.lda_matrix(), **cfg).cuda().eval()
    feats, _ = synthetic.utterances(250, np.random.RandomState(4321))
    src, mask = pad_to_longest(feats)
    batch = (None, torch.from_numpy(src).pin_memory(), torch.from_numpy(mask).pin_memory(), None, None)
    opt = types.SimpleNamespace(use_gpu=True, beam_size=10, max_token_seq_len=100, nbest=1, force_full_length=True, use_graph=False)
    translate_batch(model, batch, opt, None)
torch.cuda.synchronize()
print("done", what)
