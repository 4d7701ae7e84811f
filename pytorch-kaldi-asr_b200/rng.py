"""Dropout bookkeeping shared by all modules of one model.

The reference draws 25 independent nn.Dropout masks per forward from torch's global generator (47 % of
its CPU step).  Here a mask bit is a pure function philox(seed, site, step, element): kernels regenerate it in the
backward pass instead of storing it, and tests can materialise it (ops.dropout_keep_mask) to drive the CPU oracle
with the very same masks.  `step` lives on the device so that a captured CUDA graph sees a fresh value on replay.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib as L
from .ops import Drop


class DropoutState:
    def __init__(self, seed: int = 0):
        self.seed = int(seed)
        self.sites: Dict[str, int] = {}
        self._step: Dict[str, torch.Tensor] = {}

    def site(self, name: str) -> int:
        """Register (or look up) a dropout call site; ids are dense and assigned in registration order."""
        if name not in self.sites:
            self.sites[name] = len(self.sites) + 1
        return self.sites[name]

    def step_tensor(self, device) -> torch.Tensor:
        key = str(device)
        if key not in self._step:
            self._step[key] = torch.zeros(1, dtype=torch.int64, device=device)
        return self._step[key]

    def tick(self, device):
        """Advance the step counter on the device (stream ordered, graph capturable)."""
        if torch.device(device).type != "cuda":
            raise RuntimeError("libpka_b200 ops run on CUDA (sm_100a) tensors only; got device %s. "
                               "There is no CPU fallback." % (device,))
        t = self.step_tensor(device)
        L.check(L.lib().pka_counter_inc(L.ptr(t), L.stream_ptr()), "counter_inc")

    def make(self, p: float, site_id: int, device, training: bool) -> Optional[Drop]:
        if not training or p <= 0.0:
            return None
        return Drop(p, site_id, self.seed, self.step_tensor(device))


GLOBAL = DropoutState(seed=0)      # used by modules constructed stand-alone (outside a Transformer)
