"""Batch assembly -- the producer side of the hot path's input contract (reference U/instances_handler.py:118-139).

Only `pad_to_longest` is on the path (it fixes layout and mask polarity: 1 = real, 0 = pad, trailing padding only);
the reference's vocabulary/text helpers are data preparation and out of scope."""
from __future__ import annotations

import numpy as np

from . import constants


def pad_to_longest(instances):
    """list of 1-D (labels) or 2-D (frames x dim) arrays -> (stacked array padded with PAD, uint8 mask [n, max_len])."""
    lengths = [len(x) for x in instances]
    longest = max(lengths)
    first = np.asarray(instances[0])
    if first.ndim not in (1, 2):
        raise ValueError("undefined padding shape: instances must be 1-D or 2-D arrays")
    shape = (len(instances), longest) + tuple(first.shape[1:])
    data = np.full(shape, constants.PAD, dtype=first.dtype)
    mask = np.zeros((len(instances), longest), dtype=np.uint8)
    for i, (x, n) in enumerate(zip(instances, lengths)):
        data[i, :n] = x
        mask[i, :n] = 1
    return data, mask
