"""Restatement of the reference batched beam search (L/decode.py:22-107).  TEST INFRASTRUCTURE ONLY.

Like the reference it keeps NO cache: every step re-runs the whole decoder over the full prefixes of all live
hypotheses (L/decode.py:81-87) and advances one host-side lattice per utterance.  That is deliberate -- the oracle
is the slow, obviously-faithful version the KV-cached CUDA decoder is compared with.
"""
from __future__ import annotations

import numpy as np
import torch

from . import acoustic_model as am
from .lattice import BeamLattice


@torch.no_grad()
def translate_batch(sd, cfg, src_seq, src_pad_mask, beam_size: int, max_token_seq_len: int, nbest: int,
                    force_full_length: bool = False, return_lattices: bool = False):
    """-> (hyps[B][<=nbest][tokens], weights[B][beam]) exactly as L/decode.py:100-107.

    `force_full_length` is the benchmark stress variant of SURVEY.md 8d: the EOS log-prob is pushed to -1e30 after
    the log-softmax so no hypothesis ever ends and every utterance runs all `max_token_seq_len` steps.
    """
    off = am.DropoutPlan("off")                                   # model.eval(), L/decode.py:23
    src = torch.as_tensor(src_seq, dtype=torch.float32)
    mask = torch.as_tensor(src_pad_mask, dtype=torch.uint8)
    src, mask = am.fold_frames(src, mask, cfg["src_fold"])
    enc_out = am.encode(sd, cfg, src, mask, off)                  # once per batch, L/decode.py:46-48
    n_utt = src.shape[0]
    lattices = [BeamLattice(max_token_seq_len, beam_size) for _ in range(n_utt)]
    n_steps = 0

    for _ in range(max_token_seq_len):
        prefixes, owner = [], []
        for u, lat in enumerate(lattices):
            if not lat.done:
                seqs, _ = lat.get_results("active")
                prefixes += seqs
                owner += [u] * len(seqs)
        if not prefixes:
            break
        n_steps += 1
        tgt = torch.tensor(prefixes, dtype=torch.long)
        ones = torch.ones(tgt.shape, dtype=torch.uint8)          # all-real target mask: only the band restricts
        idx = torch.tensor(owner, dtype=torch.long)
        logits = am.decoder(sd, cfg, tgt, ones, mask.index_select(0, idx), enc_out.index_select(0, idx), off)
        word_lk = torch.log_softmax(logits[:, -1, :], dim=1).numpy()      # prob_projection, L/decode.py:87,143
        if force_full_length:
            word_lk = word_lk.copy()
            word_lk[:, am.EOS] = -1e30
        end = 0
        for lat in lattices:
            if lat.done:
                continue
            start, end = end, end + lat.num_curr_active
            lat.advance(word_lk[start:end])

    hyps, weights = [], []
    for lat in lattices:
        r, w = lat.get_results("all")
        hyps.append(r[:nbest])
        weights.append(w)
    if return_lattices:
        return hyps, weights, lattices, n_steps
    return hyps, weights
