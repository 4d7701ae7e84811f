"""numpy restatement of per-utterance cepstral mean/variance normalisation.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  In the reference CMVN is not Python at all: it is an optional call to the external Kaldi binary
`apply-cmvn` (P/run.sh:26,37-42, `cmvn=false` by default), which is neither vendored under /root/reference nor present
in this image, and no test or fixture of the reference pins its arithmetic.  What is restated here is Kaldi's published
behaviour for `apply-cmvn --utt2spk=... --norm-vars={false,true}` with per-utterance statistics (kaldi
src/transform/cmvn.cc, ApplyCmvn): mean is the plain average over the utterance's frames, variance is the *population*
variance E[x^2]-mean^2 floored at 1e-20, features become (x-mean)[/sqrt(var)].  Padded frames (beyond `length`) stay 0,
because the reference pads *after* feature extraction (U/instances_handler.py:118-139).

What freezes it (without pinning it against Kaldi): tests/golden/cmvn_hand_computed.json -- known-answer vectors worked
out by hand from that formula (perfect-square variances, an all-equal feature hitting the 1e-20 floor, N = 1 and N = 0
utterances, a padding frame), held by tests/test_oracle_vs_golden.py for this file and by tests/test_gpu_ops.py for the
CUDA front-end.
"""
from __future__ import annotations

import numpy as np


def apply_cmvn(feats: np.ndarray, lengths: np.ndarray, norm_vars: bool = False) -> np.ndarray:
    """feats f32[B,T,F] zero-padded, lengths int[B] -> normalised f32[B,T,F] (padding stays zero)."""
    out = np.zeros_like(feats, dtype=np.float32)
    for b, n in enumerate(np.asarray(lengths)):
        n = int(n)
        if n <= 0:
            continue
        x = feats[b, :n].astype(np.float64)
        mean = x.sum(axis=0) / n
        y = x - mean
        if norm_vars:
            var = (x * x).sum(axis=0) / n - mean * mean
            var = np.maximum(var, 1e-20)
            y = y / np.sqrt(var)
        out[b, :n] = y.astype(np.float32)
    return out
