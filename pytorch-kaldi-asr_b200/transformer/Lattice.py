"""`Lattice` with the reference's host API (T/Lattice.py:7-107) on top of the device-resident lattice kernel.

The batched decoder (decode.py) never goes through this class -- it keeps all utterances' lattices on the device and
advances them in one launch per step.  This single-utterance wrapper exists so that code written against the
reference's `Lattice(max_length, beam_size).advance(weights)` keeps working, and so that the one known-answer vector
the reference ships (its `main()` demo) can be checked against the very kernel the decoder uses.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib as L
from ..utils import constants


class Lattice(object):
    def __init__(self, max_length, beam_size, device="cuda"):
        self.max_length, self.beam_size = max_length, beam_size
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("Lattice runs on the GPU lattice kernel; there is no CPU path")
        self.curr_length = 0
        self.done = False
        self.num_curr_active = 1
        self._V = None
        E = self._E = 1 + beam_size * (max_length + 1)
        i32 = dict(device=self.device, dtype=torch.int32)
        self._prev = torch.full((1, E), -1, **i32)
        self._word = torch.zeros(1, E, **i32)
        self._word[0, 0] = constants.BOS
        self._depth = torch.zeros(1, E, **i32)
        self._weight = torch.zeros(1, E, device=self.device, dtype=torch.float64)
        self._n_edges = torch.ones(1, **i32)
        self._beam = torch.zeros(1, beam_size, **i32)
        self._beam_count = torch.ones(1, **i32)
        self._slot_edge = torch.zeros(1, beam_size, **i32)
        self._slot_active = torch.zeros(1, beam_size, **i32)
        self._slot_active[0, 0] = 1
        self._len = torch.zeros(1, **i32)
        self._done = torch.zeros(1, **i32)
        self._not_done = torch.ones(1, **i32)

    def advance(self, weights):
        """`weights`: [n_active, V] log-probabilities of the active hypotheses in beam order (numpy or tensor)."""
        if self.done:
            print('[WARNING] decode already finish!')
            return True
        w = torch.as_tensor(np.asarray(weights), dtype=torch.float32).to(self.device)
        V = w.shape[1]
        rows = torch.zeros(self.beam_size, V, device=self.device, dtype=torch.float32)
        rows[: w.shape[0]] = w
        d = L.BeamDesc()
        d.n_utt, d.beam, d.V, d.max_edges, d.max_len = 1, self.beam_size, V, self._E, self.max_length
        d.eos, d.force_full_length, d.inputs_are_logprobs = constants.EOS, 0, 1
        L.check(L.lib().pka_beam_advance(C.byref(d), L.ptr(rows), L.ptr(self._prev), L.ptr(self._word), L.ptr(self._depth),
                                         L.ptr(self._weight), L.ptr(self._n_edges), L.ptr(self._beam), L.ptr(self._beam_count),
                                         L.ptr(self._slot_edge), L.ptr(self._slot_active), L.ptr(self._len), L.ptr(self._done),
                                         L.ptr(self._not_done), L.stream_ptr()), "beam_advance")
        self.curr_length = int(self._len.item())
        self.done = bool(self._done.item())
        self.num_curr_active = int(self._slot_active.sum().item())
        return self.done

    # ---- read-out, same views as the reference ------------------------------------------------------------------
    @property
    def edges(self):
        n = int(self._n_edges.item())
        prev, word, wt = self._prev[0, :n].tolist(), self._word[0, :n].tolist(), self._weight[0, :n].tolist()
        return [[p, w, s] for p, w, s in zip(prev, word, wt)]

    @property
    def curr_edge_index(self):
        return self._beam[0, : int(self._beam_count.item())].tolist()

    def get_active_edge(self, edge_index):
        word = self._word[0].tolist()
        return [e for e in edge_index if word[e] != constants.EOS]

    def get_end_edge(self, edge_index):
        word = self._word[0].tolist()
        return [e for e in edge_index if word[e] == constants.EOS]

    def get_sequence(self, index):
        prev, word = self._prev[0].tolist(), self._word[0].tolist()
        out = []
        while index > -1:
            out.append(word[index])
            index = prev[index]
        return out[::-1]

    def get_results(self, mode='all'):
        cur = self.curr_edge_index
        ids = {'all': cur, 'active': self.get_active_edge(cur), 'end': self.get_end_edge(cur)}[mode]
        wt = self._weight[0].tolist()
        return [self.get_sequence(e) for e in ids], [wt[e] for e in ids]
