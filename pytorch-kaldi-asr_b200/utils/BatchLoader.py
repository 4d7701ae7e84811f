"""Batch loader feeding the hot path: same constructor, iteration protocol and output tuple as
the reference's `BatchLoader` (U/BatchLoader.py:9-107),

    for key, src_seq, src_pad_mask, tgt_seq, tgt_pad_mask in BatchLoader(triples, batch_size, pre_load, print_info, mode):

with `triples = [(key, rxfilename, label array), ...]` (L/train.py:45-54), `mode` 'drop' (training: the incomplete
last batch is skipped) or 'all' (evaluation / decoding: it is returned), a fresh shuffle at every `iter()` drawn from
the global `random` module exactly like the reference (same seed -> same batches, tests/test_loader.py checks this
against batches produced by the reference class itself).

What is different underneath, because the consumer is a ~1.5 ms GPU step rather than a 2 s one:

* pre-loaded features live once, un-padded, in one packed frame store ([sum T_i, D] fp32 + offsets); an epoch shuffle
  permutes an index list (the reference re-zips and re-materialises the whole padded set as new numpy arrays on every
  `iter()`), and a batch is one pass of row-range copies into its padded buffer;
* `next_into(alloc)` lets the consumer supply the destination buffers -- `train._Prefetcher` hands out its pinned
  staging memory, so a batch is written exactly once on the host and leaves through an asynchronous H2D copy;
* `pad_to='dataset'` (what the reference does when pre-loading: every batch has the shape of the longest utterance of
  the set, which is also what a CUDA-graph replay needs) or `'batch'` (what it does when streaming), independent of
  `pre_load`; `pad_multiple` rounds the frame axis up; `bucket=K` sorts by length inside windows of K batches so that
  'batch' padding wastes little; `pad_to='bucket'` pads a batch to `synthetic.bucket_length` (longest + 16 frames of
  context, rounded to 64, at most the set length) and its labels to a multiple of 8: a handful of shapes, one CUDA
  graph each, with results on real frames equal to whole-set padding (tests/test_gpu_bucketed.py);
* the feature gather of a pre-loaded batch runs in the native library (`pka_host_pack_batch`; `pack_threads` > 1
  splits the rows over host threads -- measured: the copy is memory-bound and one thread is fastest on the 8-core
  build container -- and 0 selects the numpy loop, which the tests hold the native path to bit for bit);
* `shard=(rank, world)` deals the epoch's batches round-robin to data-parallel ranks (needs `seed` so every rank draws
  the same permutation; in 'drop' mode every rank gets the same number of batches so collectives line up);
* `read_ahead=n` assembles up to n batches on a background thread when streaming from disk.

Errors raise `ValueError` where the reference prints `[ERROR]` and exits.
"""
from __future__ import annotations

import queue
import random
import threading
import time
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import constants, kaldi_ark

__all__ = ["BatchLoader"]

_DTYPES = (np.float32, np.uint8, np.int64, np.uint8)


def _numpy_alloc(specs):
    return [np.empty(shape, dtype=dt) for shape, dt in specs]


class BatchLoader:
    def __init__(self, trainning_triples, batch_size, pre_load=True, print_info=True, mode='drop', *,
                 pad_to: Optional[str] = None, pad_multiple: int = 1, bucket: int = 0, seed: Optional[int] = None,
                 shard: Tuple[int, int] = (0, 1), read_ahead: int = 0, reader: Optional[Callable] = None,
                 pack_threads: int = 1):
        if mode not in ('all', 'drop'):
            raise ValueError('[ERROR] mode of BatchLoader can only be [all] or [drop]')
        if batch_size < 1:
            raise ValueError('[ERROR] batch_size must be positive')
        if pad_to not in (None, 'dataset', 'batch', 'bucket'):
            raise ValueError("[ERROR] pad_to can only be 'dataset', 'batch' or 'bucket'")
        rank, world = shard
        if not 0 <= rank < world:
            raise ValueError('[ERROR] shard must be (rank, world) with 0 <= rank < world')
        if world > 1 and seed is None:
            raise ValueError('[ERROR] a sharded BatchLoader needs a seed (all ranks must draw the same permutation)')
        self.batch_size = int(batch_size)
        self.pre_load = bool(pre_load)
        self.print_info = bool(print_info)
        self.mode = mode
        self.pad_to = pad_to or ('dataset' if pre_load else 'batch')
        self.pad_multiple = max(1, int(pad_multiple))
        self.bucket = int(bucket)
        self.shard = (rank, world)
        self.read_ahead = int(read_ahead)
        self.pack_threads = int(pack_threads)
        self._rng = random if seed is None else random.Random(seed)
        self._read = reader or kaldi_ark.read_mat

        self.keys: List[str] = [t[0] for t in trainning_triples]
        self._src: List = [t[1] for t in trainning_triples]             # rxfilenames (or matrices given directly)
        self._tgt: List[np.ndarray] = [np.asarray(t[2], dtype=np.int64).reshape(-1) for t in trainning_triples]
        n = len(self.keys)
        self.num_batch = n // self.batch_size
        self._order = list(range(n))
        self.curr_iter = 0
        self._plan: List[Sequence[int]] = []
        self._producer = None

        self._tgt_len = np.array([len(t) for t in self._tgt], dtype=np.int64)
        self._max_tgt = int(self._tgt_len.max()) if n else 0
        self._frames = None           # packed store (pre_load)
        self._offsets = None
        self._src_len = None
        self.feat_dim = None
        if self.pre_load:
            self._preload()
            if self.print_info:
                print('[INFO] data preloaded.')
        elif self.pad_to in ('dataset', 'bucket') or self.bucket:
            raise ValueError("[ERROR] pad_to='dataset' and length bucketing need utterance lengths: use pre_load=True")
        if self.print_info:
            print('[INFO] loader initialized. data size:{}, batch_size:{}, iter per epoch:{}.'
                  .format(n, self.batch_size, self.num_batch))

    # ------------------------------------------------------------------------------------------------ storage
    def _load_one(self, i: int) -> np.ndarray:
        src = self._src[i]
        mat = self._read(src) if isinstance(src, str) else np.asarray(src)
        if mat.ndim != 2:
            raise ValueError('[ERROR] utterance {} is not a frames x dim matrix'.format(self.keys[i]))
        return mat

    def _preload(self):
        mats = [self._load_one(i) for i in range(len(self.keys))]
        lens = np.array([m.shape[0] for m in mats], dtype=np.int64)
        self._src_len = lens
        self._offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        if mats:
            self.feat_dim = int(mats[0].shape[1])
            if any(m.shape[1] != self.feat_dim for m in mats):
                raise ValueError('[ERROR] utterances with different feature dimensions')
            self._frames = np.empty((int(self._offsets[-1]), self.feat_dim), dtype=np.float32)
            for m, o in zip(mats, self._offsets[:-1]):
                self._frames[o:o + m.shape[0]] = m
        self._src = [None] * len(mats)

    def _utt(self, i: int) -> np.ndarray:
        if self._frames is not None:
            return self._frames[self._offsets[i]:self._offsets[i + 1]]
        return self._load_one(i)

    def _round_up(self, t: int) -> int:
        m = self.pad_multiple
        return (t + m - 1) // m * m

    # ------------------------------------------------------------------------------------------------ resume
    def get_rng_state(self):
        """State of the epoch-shuffle generator plus the current (cumulatively shuffled) order -- what a checkpoint needs
        for the next epoch's permutation to be the one an uninterrupted run would draw."""
        state = self._rng.getstate() if hasattr(self._rng, 'getstate') else None
        return dict(rng=state, order=list(self._order))

    def set_rng_state(self, state):
        if state.get('rng') is not None and hasattr(self._rng, 'setstate'):
            rng = state['rng']
            self._rng.setstate((rng[0], tuple(rng[1]), rng[2]) if isinstance(rng, (list, tuple)) else rng)
        if state.get('order') is not None and len(state['order']) == len(self._order):
            self._order[:] = list(state['order'])

    # ------------------------------------------------------------------------------------------------ epoch plan
    def __len__(self):
        rank, world = self.shard
        n_full, tail = self.num_batch, (len(self.keys) % self.batch_size != 0)
        total = n_full + (1 if (self.mode == 'all' and tail) else 0)
        if world == 1:
            return total
        return n_full // world if self.mode == 'drop' else len(range(rank, total, world))

    def __iter__(self):
        self._stop_producer()
        self.curr_iter = 0
        self._rng.shuffle(self._order)                      # cumulative, like the reference's in-place re-shuffle
        order = self._order
        bs = self.batch_size
        batches = [order[i * bs:(i + 1) * bs] for i in range(self.num_batch)]
        if self.bucket > 0 and self._src_len is not None:
            win = self.bucket * bs
            batches = []
            for w in range(0, self.num_batch * bs, win):
                chunk = sorted(order[w:min(w + win, self.num_batch * bs)], key=lambda i: int(self._src_len[i]))
                group = [chunk[i:i + bs] for i in range(0, len(chunk), bs)]
                self._rng.shuffle(group)
                batches.extend(group)
        tail = order[self.num_batch * bs:]
        rank, world = self.shard
        if world > 1:
            usable = len(batches) // world * world
            mine = batches[rank:usable:world]
            # 'all' (evaluation, no collectives): the left-over full batches and the tail are dealt on round-robin too
            extra = batches[usable:] + ([tail] if tail else [])
            self._plan_all_extra = [b for j, b in enumerate(extra, usable) if j % world == rank]
            self._plan = mine
        else:
            self._plan = batches
            self._plan_all_extra = [tail] if tail else []
        if self.print_info:
            print('[INFO] script list is shuffled')
        if self.read_ahead > 0 and not self.pre_load:
            self._start_producer()
        return self

    def _next_indices(self) -> Sequence[int]:
        """Index list of the next batch under the *current* `mode` (the reference also decides this per call)."""
        k = self.curr_iter
        if k < len(self._plan):
            idx = self._plan[k]
        elif self.mode == 'all' and k - len(self._plan) < len(self._plan_all_extra):
            idx = self._plan_all_extra[k - len(self._plan)]
        else:
            raise StopIteration()
        self.curr_iter += 1
        return idx

    # ------------------------------------------------------------------------------------------------ assembly
    def _assemble(self, idx: Sequence[int], alloc) -> tuple:
        mats = None
        if self._frames is None:
            mats = [self._load_one(i) for i in idx]
            t_max = max(m.shape[0] for m in mats)
            dim = mats[0].shape[1]
        else:
            dim = self.feat_dim
            t_max = int(self._src_len.max()) if self.pad_to == 'dataset' else int(self._src_len[list(idx)].max())
        t_pad = self._round_up(t_max)
        l_pad = self._max_tgt if self.pad_to == 'dataset' else int(self._tgt_len[list(idx)].max())
        if self.pad_to == 'bucket':
            # a few padded shapes (one CUDA graph each, train.GraphedTrainStep) instead of one per batch: frames to the
            # next multiple of 64 that leaves the TDNN stack's 16 frames of right context as padding (results on real
            # frames then equal those under whole-set padding), never beyond the whole-set length; labels to 8 tokens
            from .synthetic import bucket_length
            t_pad = bucket_length(t_max, int(self._src_len.max()))
            l_pad = min(self._max_tgt, (l_pad - 1 + 7) // 8 * 8 + 1)
        b = len(idx)
        src, src_mask, tgt, tgt_mask = alloc([((b, t_pad, dim), _DTYPES[0]), ((b, t_pad), _DTYPES[1]),
                                              ((b, l_pad), _DTYPES[2]), ((b, l_pad), _DTYPES[3])])
        native = (mats is None and self.pack_threads > 0 and b > 0 and src.flags['C_CONTIGUOUS']
                  and src_mask.flags['C_CONTIGUOUS'])
        if native:
            self._native_pack(idx, t_pad, src, src_mask)
        for row, i in enumerate(idx):
            if not native:
                m = mats[row] if mats is not None else self._utt(i)
                t = m.shape[0]
                src[row, :t] = m
                src[row, t:] = constants.PAD
                src_mask[row, :t] = 1
                src_mask[row, t:] = 0
            lab = self._tgt[i]
            n = lab.shape[0]
            tgt[row, :n] = lab
            tgt[row, n:] = constants.PAD
            tgt_mask[row, :n] = 1
            tgt_mask[row, n:] = 0
        return tuple(self.keys[i] for i in idx), src, src_mask, tgt, tgt_mask

    def _native_pack(self, idx, t_pad, src, src_mask):
        import ctypes as C
        from .. import _lib as L
        order = np.ascontiguousarray(idx, dtype=np.int64)
        L.check(L.lib().pka_host_pack_batch(C.c_void_p(self._frames.ctypes.data), C.c_void_p(self._offsets.ctypes.data),
                                            C.c_int64(len(self.keys)), C.c_void_p(order.ctypes.data), len(order),
                                            int(t_pad), int(self.feat_dim), C.c_void_p(src.ctypes.data),
                                            C.c_void_p(src_mask.ctypes.data), self.pack_threads), "host_pack_batch")

    def next_into(self, alloc) -> tuple:
        """next(), with the four arrays written into buffers from `alloc([(shape, dtype) x 4]) -> [ndarray x 4]`."""
        start_time = time.time() if self.print_info else 0.0
        if self._producer is not None:
            item = self._producer_get()
            key, arrays = item[0], item[1:]
            outs = alloc([(a.shape, a.dtype) for a in arrays])
            for o, a in zip(outs, arrays):
                o[...] = a
            batch = (key,) + tuple(outs)
        else:
            batch = self._assemble(self._next_indices(), alloc)
        if self.print_info:
            print('[INFO] iter {}: data loaded. loading cost {:3.2f} seconds'.format(self.curr_iter, time.time() - start_time))
        return batch

    def __next__(self):
        return self.next_into(_numpy_alloc)

    # ------------------------------------------------------------------------------------------------ read-ahead
    def _start_producer(self):
        q: "queue.Queue" = queue.Queue(maxsize=self.read_ahead)
        stop = threading.Event()
        plan, extra = list(self._plan), list(self._plan_all_extra)

        def put(item):
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.05)
                    return True
                except queue.Full:
                    continue
            return False

        def work():
            try:
                for n, idx in enumerate(plan + extra):
                    if not put(('tail' if n >= len(plan) else 'full', self._assemble(idx, _numpy_alloc))):
                        return
                put(('end', None))
            except BaseException as exc:                      # surfaced in the consumer thread
                put(('error', exc))

        th = threading.Thread(target=work, name='BatchLoader-read-ahead', daemon=True)
        self._producer = (q, stop, th)
        th.start()

    def _producer_get(self):
        q, _, _ = self._producer
        kind, item = q.get()
        if kind == 'error':
            self._stop_producer()
            raise item
        if kind == 'end' or (kind == 'tail' and self.mode != 'all'):
            self._stop_producer()
            raise StopIteration()
        self.curr_iter += 1
        return item

    def _stop_producer(self):
        if self._producer is not None:
            q, stop, th = self._producer
            stop.set()
            th.join()
            self._producer = None

    def __del__(self):
        try:
            self._stop_producer()
        except Exception:
            pass
