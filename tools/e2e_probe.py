"""Where does the end-to-end step lose time against the resident step?  CPU-side phase timing of train_epoch's loop."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pytorch_kaldi_asr_b200 as pk
from pytorch_kaldi_asr_b200 import train as T
from pytorch_kaldi_asr_b200.utils import synthetic
pk.set_compute_mode("bf16")
cfg = dict(n_src_dim=40, n_tgt_vocab=53, encoder_max_len=500, decoder_max_len=100, src_fold=1, encoder_sub_sequence=(-100, 0),
           decoder_sub_sequence=(-10, 0), en_layers=3, de_layers=3, n_head=2, en_d_model=256, de_d_model=128, d_k=64, d_v=64,
           en_dropout=0.35, de_dropout=0.35, tdnn_contexts=[[-1, 0, 1], [-1, 0, 1], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3]])
torch.manual_seed(0)
model = pk.Transformer(lda_mat=synthetic.lda_matrix(), **cfg).cuda()
opt = pk.ScheduledOptim(pk.FusedAdam(model.parameters()), 1e-3, 25000)
pool = synthetic.batches(8, 32, seed=1234, pad_to="set")
pin = [(None,) + tuple(torch.as_tensor(np.ascontiguousarray(x)).to(dt).pin_memory() for x, dt in zip(b[1:], (torch.float32, torch.uint8, torch.int64, torch.uint8))) for b in pool]
g = pk.GraphedTrainStep(model, opt, pool[0])
class Loader(list):
    mode = "drop"
def run(n, sync):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pk.train_epoch(model, Loader([pin[i % 8] for i in range(n)]), None, mode="train", optimizer=opt, graphed=g, sync_every_step=sync)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
for sync in (True, False, True, False):
    run(5, sync)
    print("sync_every_step=%s: %.3f ms/step" % (sync, run(60, sync)))
# resident replay for reference
dev = [T._to_device(b, "cuda") for b in pool]
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(60):
    g.load(*dev[i % 8]); g.graph.replay()
torch.cuda.synchronize(); print("resident: %.3f ms/step" % ((time.perf_counter() - t0) / 60 * 1e3))
# CPU cost of one loop iteration without GPU waits
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); run(60, True); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
