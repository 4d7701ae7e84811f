"""Per-parameter comparison of one backward pass with and without deferred reductions (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_kaldi_asr_b200 as pk
from pytorch_kaldi_asr_b200 import ops
from pytorch_kaldi_asr_b200.utils import synthetic
from oracle import acoustic_model as am

def grads(defer, dropout, hints=True):
    cfg = am.example_config(en_dropout=dropout, de_dropout=dropout)
    lda = synthetic.lda_matrix()
    sd = am.init_state_dict(cfg, lda, seed=0)
    batch = synthetic.batches(1, 5, seed=31)[0]
    model = pk.Transformer(lda_mat=lda, seed=3, **{k: v for k, v in cfg.items() if k != "encoder_type"})
    model.load_state_dict(sd)
    model = model.cuda().train()
    opt = pk.FusedAdam(model.parameters())
    opt.zero_grad()
    ops.DEFER_ENABLED = defer
    pk.set_compute_mode("bf16")
    src, smask, tgt, tmask = pk.train._to_device(batch, "cuda")
    tgt_in, goal, tmask_in = ops.split_targets(tgt, tmask)
    pred = model(src, smask, tgt_in, tmask_in)
    loss, _ = ops.cross_entropy_sum(pred.view(-1, pred.size(-1)), goal.view(-1), False)
    loss.backward()
    torch.cuda.synchronize()
    pk.set_compute_mode("fp32")
    return {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}

for dropout in (0.0, 0.35):
    g0, g1 = grads(False, dropout), grads(True, dropout)
    print("dropout", dropout)
    for k in g0:
        rel = float((g0[k] - g1[k]).abs().max() / g0[k].abs().max().clamp_min(1e-20))
        if rel > 1e-5:
            print("  %-60s rel %.3e  |g0| %.3e |g1| %.3e" % (k, rel, float(g0[k].abs().max()), float(g1[k].abs().max())))
