"""Kernel-time breakdown of the TIMIT training step as it runs inside the CUDA graph (torch.profiler / CUPTI; hot-L2
durations, unlike the cold-cache serialised ncu launch list).  usage: python tools/profile_step.py [bf16|fp32] [replays]"""
import collections, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_kaldi_asr_b200 as pk
from pytorch_kaldi_asr_b200.utils import synthetic
from torch.profiler import profile, ProfilerActivity
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
replays = int(sys.argv[2]) if len(sys.argv) > 2 else 5
pk.set_compute_mode(mode)
cfg = dict(n_src_dim=40, n_tgt_vocab=53, encoder_max_len=500, decoder_max_len=100, src_fold=1, encoder_sub_sequence=(-100, 0),
           decoder_sub_sequence=(-10, 0), en_layers=3, de_layers=3, n_head=2, en_d_model=256, de_d_model=128, d_k=64, d_v=64,
           en_dropout=0.35, de_dropout=0.35, tdnn_contexts=[[-1, 0, 1], [-1, 0, 1], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3]])
torch.manual_seed(0)
model = pk.Transformer(lda_mat=synthetic.lda_matrix(), **cfg).cuda()
opt = pk.ScheduledOptim(pk.FusedAdam(model.parameters()), 1e-3, 25000)
batch = synthetic.batches(1, 32, seed=1234, pad_to="set")[0]
g = pk.GraphedTrainStep(model, opt, batch)
for _ in range(3):
    g.graph.replay()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    g.graph.replay()
e.record()
torch.cuda.synchronize()
print("graph replay: %.3f ms/step" % (s.elapsed_time(e) / 20))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(replays):
        g.graph.replay()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"\(.*", "", ev.name)[:70]
        agg[name][0] += 1
        agg[name][1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
tot = sum(v[1] for v in agg.values())
print("kernel time per step: %.1f us in %d launches" % (tot / replays, sum(v[0] for v in agg.values()) / replays))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%9.1f us/step %5.1f launches %7.2f us each %5.1f%%  %s" % (v[1] / replays, v[0] / replays, v[1] / v[0], 100 * v[1] / tot, k))
