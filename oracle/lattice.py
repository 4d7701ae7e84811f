"""numpy restatement of the reference beam lattice (T/Lattice.py:7-107).  TEST INFRASTRUCTURE ONLY.

State is kept as three parallel columns (back-pointer, word, cumulative weight) instead of the reference's list of
3-element lists; `edges` rebuilds the reference's view for comparisons.  Cumulative weights are float64, as in the
reference (np.append / python floats promote the fp32 log-probs, T/Lattice.py:45-54).

Tie rule: the reference ranks with `np.argsort(-w)` (introsort, unstable -> tie order unspecified, T/Lattice.py:56).
This restatement uses a *stable* sort, i.e. ties resolve to the lowest flat candidate index; that is the rule the CUDA
top-k implements.  Parity tests assert that their inputs contain no exact ties among the kept candidates.
"""
from __future__ import annotations

import numpy as np

PAD, UNK, BOS, EOS = 0, 1, 2, 3


class BeamLattice:
    def __init__(self, max_length: int, beam_size: int):
        self.max_length = max_length
        self.beam_size = beam_size
        self.prev = [-1]                 # back-pointer of every edge ever created
        self.word = [BOS]
        self.weight = [0.0]              # float64 cumulative log-prob
        self.beam = [0]                  # edge ids currently in the beam, best first (finished edges stay in it)
        self.curr_length = 0
        self.num_curr_active = 1
        self.done = False
        self.min_gap = float("inf")

    # -- views -------------------------------------------------------------------------------------------
    @property
    def edges(self):
        return [[p, w, s] for p, w, s in zip(self.prev, self.word, self.weight)]

    def active(self):
        return [e for e in self.beam if self.word[e] != EOS]

    def finished(self):
        return [e for e in self.beam if self.word[e] == EOS]

    # -- one step (T/Lattice.py:35-81) -------------------------------------------------------------------
    def advance(self, log_probs: np.ndarray) -> bool:
        """`log_probs` is [n_active, V] for the active edges in beam order."""
        live = self.active()
        if not live:
            return True
        n_words = log_probs.shape[1]
        if len(self.prev) == 1:                       # first step: only the BOS row is used (T/Lattice.py:41-42)
            cand = np.asarray(log_probs[0], dtype=np.float64).copy()
        else:
            base = np.asarray([self.weight[e] for e in live], dtype=np.float64)
            cand = np.asarray(log_probs).reshape(-1).astype(np.float64) + np.repeat(base, n_words)
        parents = np.repeat(np.asarray(live), n_words)
        n_live_cand = len(parents)
        fin = self.finished()
        cand = np.concatenate([cand, np.asarray([self.weight[e] for e in fin], dtype=np.float64)])
        full_order = np.argsort(-cand, kind="stable")
        order = full_order[: self.beam_size]
        # smallest gap between neighbours among the kept candidates and the first rejected one: a parity test may
        # only demand identical tokens when this exceeds the fp32 error of the compared implementations.
        top = cand[full_order[: self.beam_size + 1]]
        if len(top) > 1:
            self.min_gap = min(self.min_gap, float(np.min(top[:-1] - top[1:])))

        new_beam = []
        for idx in order:
            if idx < n_live_cand:                     # expansion of a live hypothesis -> new edge
                self.prev.append(int(parents[idx]))
                self.word.append(int(idx % n_words))
                self.weight.append(float(cand[idx]))
                new_beam.append(len(self.prev) - 1)
            else:                                     # a finished hypothesis keeps its old edge id
                new_beam.append(fin[idx - n_live_cand])
        self.beam = new_beam
        self.curr_length += 1
        self.num_curr_active = len(self.active())
        if self.num_curr_active == 0 or self.curr_length > self.max_length:
            self.done = True
        return self.done

    # -- read-out (T/Lattice.py:84-107) ------------------------------------------------------------------
    def sequence(self, edge: int):
        out = []
        while edge > -1:
            out.append(self.word[edge])
            edge = self.prev[edge]
        return out[::-1]

    def get_results(self, mode: str = "all"):
        ids = {"all": self.beam, "active": self.active(), "end": self.finished()}[mode]
        return [self.sequence(e) for e in ids], [self.weight[e] for e in ids]


def demo_known_answer():
    """The only known-answer vector the reference ships: T/Lattice.py:109-130 (vocab 7, beam 3, max_len 10)."""
    lat = BeamLattice(10, 3)
    lat.advance(np.array([[-99, -99, -99, -4, -3, -2, -1]] * 3))
    lat.advance(np.array([[-99, -99, -99, -1.5, -2, -3, -4],
                          [-99, -99, -99, -1.5, -3, -4, -2],
                          [-99, -99, -99, -1.5, -4, -3, -2]]))
    lat.advance(np.array([[-99, -99, -99, -1.5, -2, -3, -4]]))
    results, weights = lat.get_results()
    return lat.done, results, weights, lat.edges
