#!/usr/bin/env python
"""Generate the checkpoint fixtures by running the REAL reference in place (build container only):

    python tests/golden/make_checkpoint_golden.py

* ref_checkpoint_tiny.torch -- a checkpoint exactly as L/initialize_model.py:90-95 writes it: the reference's own
  `Transformer` object pickled whole, its argparse Namespace as 'model_options', 'epoch': 0.
* checkpoint_tiny.npz -- that model's state dict as plain arrays, and the running averages the reference's
  `scale_dict` / `add_dict` (L/train.py:276-303) produce over four seeded perturbations of it.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
R = os.environ.get("PKA_REFERENCE", "/root/reference")
N_AVG = 4


def perturbed(sd, i):
    """The i-th 'epoch' of the averaging fixture: deterministic, different per tensor and per epoch."""
    out = {}
    for j, (k, v) in enumerate(sd.items()):
        noise = np.random.RandomState(1000 * i + j).randn(*v.shape).astype(np.float32)
        out[k] = v + 0.05 * torch.from_numpy(noise).reshape(v.shape)
    return out


def main():
    sys.path[:0] = [R + "/pytorch", R + "/project/attention-transformer-timit/local/pytorch",
                    R + "/project/attention-transformer-timit/local"]
    sys.modules.setdefault("kaldi_io", types.ModuleType("kaldi_io"))
    import train as ref_train
    from transformer.Models import Transformer
    torch.manual_seed(0)
    opt = argparse.Namespace(read_feats_scp_file="data/train/feats.scp", lda_mat_file="exp/lda.mat",
                             read_vocab_file="exp/vocab", encoder_max_len=20, decoder_max_len=12, src_fold=1,
                             encoder_sub_sequence=(-100, 0), decoder_sub_sequence=(-3, 0), en_layers=2, de_layers=1,
                             n_head=2, en_d_model=16, de_d_model=16, d_k=8, d_v=8, en_dropout=0.1, de_dropout=0.1,
                             save_model_file="exp/model.init", tdnn_contexts=[[-1, 0, 1], [-3, 0, 3]], src_dim=4,
                             tgt_vocab_dim=9)
    lda = (np.random.RandomState(0).randn(20, 21) * 0.1).astype(np.float32)
    model = Transformer(opt.src_dim, opt.tgt_vocab_dim, lda_mat=lda, encoder_max_len=opt.encoder_max_len,
                        decoder_max_len=opt.decoder_max_len, src_fold=opt.src_fold, encoder_sub_sequence=(-100, 0),
                        decoder_sub_sequence=opt.decoder_sub_sequence, en_layers=opt.en_layers, de_layers=opt.de_layers,
                        n_head=opt.n_head, en_d_model=opt.en_d_model, de_d_model=opt.de_d_model, d_k=opt.d_k, d_v=opt.d_v,
                        en_dropout=opt.en_dropout, de_dropout=opt.de_dropout, tdnn_contexts=opt.tdnn_contexts)
    torch.save({'model': model, 'model_options': opt, 'epoch': 0}, os.path.join(HERE, "ref_checkpoint_tiny.torch"))
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    out = {"sd." + k: v.numpy() for k, v in sd.items()}
    out["lda_mat"] = lda
    # the loop of L/train.py:293-303, on state dicts
    avg = None
    for i in range(N_AVG):
        nxt = perturbed(sd, i)
        if i == 0:
            avg = nxt
        else:
            factor = 1 / (i + 1)
            avg = ref_train.add_dict(ref_train.scale_dict(avg, 1 - factor), factor, nxt)
        for k, v in avg.items():
            out["avg%d.%s" % (i + 1, k)] = v.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "checkpoint_tiny.npz"), **out)
    print("wrote ref_checkpoint_tiny.torch, checkpoint_tiny.npz (%d tensors)" % len(sd))


if __name__ == "__main__":
    main()
