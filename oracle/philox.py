"""TEST INFRASTRUCTURE ONLY: numpy restatement of the dropout keep-bit definition used by every libpka_b200 kernel.

The reference draws its dropout masks from torch's global generator (nn.Dropout, e.g. T/SubLayers.py:36, L/pytorch/
TDNN.py:39); a bit-identical stream cannot be reproduced by a kernel that regenerates masks in the backward pass, so
the product defines its own counter-based stream and the parity tests inject exactly these bits into the oracle
(oracle/acoustic_model.py: DropoutPlan("injected")).  Definition (csrc/common.cuh):

    r = philox4x32-10(key = (seed_lo, seed_hi), counter = (idx8_lo, idx8_hi ^ step_hi, site, step_lo)),  idx8 = i >> 3
    lane = 16-bit half (i & 7) of the 128-bit result (little end first: r0 low, r0 high, r1 low, ...)
    keep(i) = lane >= floor(p * 2^16)
"""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(seed: int, site: int, step: int, idx: np.ndarray) -> np.ndarray:
    idx = np.asarray(idx, dtype=np.uint64)
    k0 = np.uint64(seed & MASK)
    k1 = np.uint64((seed >> 32) & MASK)
    c0 = idx & np.uint64(MASK)
    c1 = (idx >> np.uint64(32)) ^ np.uint64((step >> 32) & MASK)
    c2 = np.full_like(idx, site & MASK)
    c3 = np.full_like(idx, step & MASK)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        h0, l0 = p0 >> np.uint64(32), p0 & np.uint64(MASK)
        h1, l1 = p1 >> np.uint64(32), p1 & np.uint64(MASK)
        c0, c1, c2, c3 = (h1 ^ c1 ^ k0) & np.uint64(MASK), l1, (h0 ^ c3 ^ k1) & np.uint64(MASK), l0
        k0 = (k0 + np.uint64(W0)) & np.uint64(MASK)
        k1 = (k1 + np.uint64(W1)) & np.uint64(MASK)
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint64)


def keep_mask(n: int, p: float, site: int, seed: int, step: int) -> np.ndarray:
    """uint8[n]: 1 = kept, for elements 0..n-1 of one dropout site at one training step."""
    if p <= 0.0:
        return np.ones(n, np.uint8)
    i = np.arange(n, dtype=np.uint64)
    r = philox4x32_10(seed, site, step, i >> np.uint64(3))
    word = r[np.arange(n), ((i & np.uint64(7)) >> np.uint64(1)).astype(np.int64)]
    lane = np.where((i & np.uint64(1)) == 0, word & np.uint64(0xFFFF), word >> np.uint64(16))
    t = p * 65536.0
    thresh = 0xFFFF if t >= 65535.0 else int(t)
    return (lane >= np.uint64(thresh)).astype(np.uint8)
