"""Seeded synthetic TIMIT-shaped data, exactly as fixed in SURVEY.md 8(d) / BASELINE.md 3 so that the B200 path and
the CPU baseline consume identical inputs (there are no datasets on the boxes).

    per utterance  T_i = clip(round(N(307, 86)), 90, 499) frames of N(0,1) fp32 features (40-dim), zero padded
    labels         [BOS] + randint(4, 52, size=clip(T_i // 8, 5, 98)) + [EOS]
    lda_mat        RandomState(0).randn(200, 201) * 0.1
"""
from __future__ import annotations

import numpy as np

from . import constants
from .instances_handler import pad_to_longest


def lda_matrix(feat_dim: int = 40, fold: int = 1, seed: int = 0) -> np.ndarray:
    s = feat_dim * fold * 5
    return (np.random.RandomState(seed).randn(s, s + 1) * 0.1).astype(np.float32)


def utterances(n: int, rng: np.random.RandomState, feat_dim: int = 40, vocab: int = 53, mean_len: float = 307.0,
               std_len: float = 86.0, min_len: int = 90, max_len: int = 499, label_div: int = 8, max_labels: int = 98):
    feats, labels = [], []
    for _ in range(n):
        t = int(np.clip(int(round(rng.normal(mean_len, std_len))), min_len, max_len))
        feats.append(rng.randn(t, feat_dim).astype(np.float32))
        n_lab = int(np.clip(t // label_div, 5, max_labels))
        body = rng.randint(4, vocab - 1, size=n_lab)
        labels.append(np.concatenate([[constants.BOS], body, [constants.EOS]]).astype(np.int64))
    return feats, labels


def batches(n_batches: int, batch_size: int, seed: int = 1234, pad_to: str = "batch", **kw):
    """-> list of (keys, src f32[B,T,F], src_mask u8[B,T], tgt i64[B,L+1], tgt_mask u8[B,L+1]).

    pad_to="batch": pad each batch to its own longest utterance (pad_to_longest).
    pad_to="set":   pad every batch to the longest utterance of the whole set, like the reference's pre-loading
                    BatchLoader (U/BatchLoader.py:33-36) -- all batches then share one shape (one CUDA graph).
    pad_to="bucket": utterances sorted by length into batches (batch order shuffled), each batch padded to
                    bucket_length(longest, set length); labels to a multiple of 8 tokens (a few shapes, one graph each)."""
    rng = np.random.RandomState(seed)
    feats, labels = utterances(n_batches * batch_size, rng, **kw)
    if pad_to == "bucket":
        t_set = max(f.shape[0] for f in feats)
        order = np.argsort([f.shape[0] for f in feats], kind="stable")
        out = []
        # the batch order is the same for every seed: data-parallel ranks (one seed each) then step through their
        # length quantiles in the same order and see similar shapes in the same step (no straggler rank)
        for b in np.random.RandomState(4321).permutation(n_batches):
            idx = order[b * batch_size:(b + 1) * batch_size]
            fs, ls = [feats[i] for i in idx], [labels[i] for i in idx]
            t_pad = bucket_length(max(f.shape[0] for f in fs), t_set)
            l_pad = (max(len(l) for l in ls) - 1 + 7) // 8 * 8 + 1
            src, smask = pad_to_longest(fs, t_pad)
            tgt, tmask = pad_to_longest(ls, l_pad)
            out.append((["utt%06d" % i for i in idx], src, smask, tgt, tmask))
        return out
    if pad_to == "set":
        src_all, smask_all = pad_to_longest(feats)
        tgt_all, tmask_all = pad_to_longest(labels)
    out = []
    for b in range(n_batches):
        sl = slice(b * batch_size, (b + 1) * batch_size)
        keys = ["utt%06d" % i for i in range(sl.start, sl.stop)]
        if pad_to == "set":
            out.append((keys, src_all[sl], smask_all[sl], tgt_all[sl], tmask_all[sl]))
        else:
            src, smask = pad_to_longest(feats[sl])
            tgt, tmask = pad_to_longest(labels[sl])
            out.append((keys, src, smask, tgt, tmask))
    return out


def bucket_length(longest: int, set_length: int, context: int = 16, quantum: int = 64) -> int:
    """Padded length of a batch whose longest utterance has `longest` frames: at least `context` pad frames behind
    every utterance (the TDNN stack's right receptive field, SURVEY 8e: the encoder output of real frames is then
    independent of any further padding), rounded up to `quantum`, never beyond the whole-set length the reference
    pads to (a batch that long sees exactly the reference's padding)."""
    return int(min((longest + context + quantum - 1) // quantum * quantum, set_length))


def real_frames(batch) -> int:
    return int(np.asarray(batch[2]).sum())
