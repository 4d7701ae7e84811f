"""Batch assembly -- the producer side of the hot path's input contract (reference U/instances_handler.py:118-139).

`pad_to_longest` is on the path (it fixes layout and mask polarity: 1 = real, 0 = pad, trailing padding only).  The
label helpers `initialize_batch_loader` needs (`read_vocab`, `add_control_words`, `apply_vocab`; U/instances_handler.py:
73-110) are here for the loader; vocabulary building and text preparation stay out of scope."""
from __future__ import annotations

import numpy as np

from . import constants


def pad_to_longest(instances, length=None):
    """list of 1-D (labels) or 2-D (frames x dim) arrays -> (stacked array padded with PAD, uint8 mask [n, max_len]).
    `length` (extension; the reference pads to the longest instance, U/instances_handler.py:118-139) pads further, to a
    bucket length >= the longest instance."""
    lengths = [len(x) for x in instances]
    longest = max(lengths)
    if length is not None:
        if length < longest:
            raise ValueError("pad_to_longest: length %d is shorter than the longest instance (%d)" % (length, longest))
        longest = int(length)
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


def read_vocab(vocab_file):
    """Symbol table 'word index' per line -> {word: index} (U/instances_handler.py:73-82)."""
    word2idx = {}
    with open(vocab_file, encoding='utf-8') as f:
        for n, line in enumerate(f, 1):
            fields = line.split()
            if not fields:
                continue
            if len(fields) < 2:
                raise ValueError('[ERROR] {} line {}: expected "word index"'.format(vocab_file, n))
            word2idx[fields[0]] = int(fields[1])
    return word2idx


def add_control_words(instances_index):
    """{key: [words]} -> {key: array([<s>, words..., </s>])}, in place like the reference (U/instances_handler.py:86-90)."""
    for key, words in instances_index.items():
        instances_index[key] = np.array([constants.BOS_WORD] + list(words) + [constants.EOS_WORD])
    return instances_index


def apply_vocab(instances, vocab_file, mode):
    """mode 'word2idx': words -> int64 indices (UNK for unknown words); 'idx2word': the inverse (UNK_WORD for unknown
    indices).  `vocab_file` may also be an already loaded {word: index} dict.  (U/instances_handler.py:94-110)"""
    word2idx = vocab_file if isinstance(vocab_file, dict) else read_vocab(vocab_file)
    if mode == 'word2idx':
        return {key: np.array([word2idx.get(w, constants.UNK) for w in words], dtype=np.int64)
                for key, words in instances.items()}
    if mode == 'idx2word':
        idx2word = {i: w for w, i in word2idx.items()}
        return {key: [idx2word.get(int(i), constants.UNK_WORD) for i in idxs] for key, idxs in instances.items()}
    raise ValueError('[ERROR] invalid mode string: {!r}'.format(mode))
