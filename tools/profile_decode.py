"""Kernel-time breakdown of beam-10 decoding (config 4 shape: 125 utterances per call, forced 100 steps)."""
import collections, os, re, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pytorch_kaldi_asr_b200 as pk
from pytorch_kaldi_asr_b200.decode import translate_batch
from pytorch_kaldi_asr_b200.utils import synthetic
from pytorch_kaldi_asr_b200.utils.instances_handler import pad_to_longest
from torch.profiler import profile, ProfilerActivity
pk.set_compute_mode("fp32")
cfg = dict(n_src_dim=40, n_tgt_vocab=53, encoder_max_len=500, decoder_max_len=100, src_fold=1, encoder_sub_sequence=(-100, 0),
           decoder_sub_sequence=(-10, 0), en_layers=3, de_layers=3, n_head=2, en_d_model=256, de_d_model=128, d_k=64, d_v=64,
           en_dropout=0.35, de_dropout=0.35, tdnn_contexts=[[-1, 0, 1], [-1, 0, 1], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3]])
torch.manual_seed(0)
model = pk.Transformer(lda_mat=synthetic.lda_matrix(), **cfg).cuda().eval()
feats, _ = synthetic.utterances(125, np.random.RandomState(4321))
src, mask = pad_to_longest(feats)
batch = (None, torch.from_numpy(src).pin_memory(), torch.from_numpy(mask).pin_memory(), None, None)
opt = types.SimpleNamespace(use_gpu=True, beam_size=10, max_token_seq_len=100, nbest=1, force_full_length=True)
translate_batch(model, batch, opt, None)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3):
    translate_batch(model, batch, opt, None)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
print("translate_batch(125 utts, 100 forced steps): %.1f ms -> %.0f utts/s, %.1f us per decoder step" % (dt * 1e3, 125 / dt, dt * 1e6 / 100))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    translate_batch(model, batch, opt, None)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        agg[re.sub(r"\(.*", "", ev.name)[:70]][0] += 1
        agg[re.sub(r"\(.*", "", ev.name)[:70]][1] += ev.device_time_total
tot = sum(v[1] for v in agg.values())
print("kernel time: %.1f ms in %d launches" % (tot / 1e3, sum(v[0] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print("%9.1f us %6d launches %7.2f us each %5.1f%%  %s" % (v[1], v[0], v[1] / v[0], 100 * v[1] / tot, k))
