"""Kernel-time breakdown of one eager training step (torch.profiler; shares only, not absolutes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_kaldi_asr_b200 as pk
from pytorch_kaldi_asr_b200.utils import synthetic
from torch.profiler import profile, ProfilerActivity
mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
pk.set_compute_mode(mode)
cfg = dict(n_src_dim=40, n_tgt_vocab=53, encoder_max_len=500, decoder_max_len=100, src_fold=1, encoder_sub_sequence=(-100, 0),
           decoder_sub_sequence=(-10, 0), en_layers=3, de_layers=3, n_head=2, en_d_model=256, de_d_model=128, d_k=64, d_v=64,
           en_dropout=0.35, de_dropout=0.35, tdnn_contexts=[[-1, 0, 1], [-1, 0, 1], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3]])
torch.manual_seed(0)
model = pk.Transformer(lda_mat=synthetic.lda_matrix(), **cfg).cuda()
opt = pk.ScheduledOptim(pk.FusedAdam(model.parameters()), 1e-3, 25000)
batch = synthetic.batches(1, 32, seed=1234, pad_to="set")[0]
class Loader(list):
    mode = "drop"
for _ in range(3):
    pk.train_epoch(model, Loader([batch]), None, mode="train", optimizer=opt)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    pk.train_epoch(model, Loader([batch]), None, mode="train", optimizer=opt)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
