"""Data-parallel training over utterance batches: one process per GPU, NCCL all-reduce of the flat gradient arena.

The reference is single-device; this is the new exchange step of 8(e).  Its loss is a SUM over tokens
(L/train.py:86-88), so gradients are all-reduced with SUM and NOT divided by the world size: N ranks x B utterances then
reproduce exactly the single-process gradient on the concatenated N*B batch.

Overlap: parameters sit in the arena in forward order, so backward completes the arena from its tail.  The arena is cut
into a few contiguous buckets; a post-accumulate hook counts finished tensors per bucket and, when a bucket is complete,
launches its all-reduce on a side stream while backward continues on the compute stream.  `finish()` (called before the
optimiser step) makes the compute stream wait for the communication stream.  Works with `gloo` on CPU tensors for tests.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of `n_items` for `rank` (decode shards utterances with no collective)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def plan_buckets(offsets: List[int], sizes: List[int], total: int, n_buckets: int):
    """Cut [0,total) into <= n_buckets contiguous ranges on tensor boundaries, roughly equal in bytes.
    Returns (bounds [(lo,hi)...] in arena order, bucket index per tensor)."""
    n_buckets = max(1, min(n_buckets, len(offsets)))
    target = total / n_buckets
    bounds, owner, lo, b = [], [], 0, 0
    for i, (off, sz) in enumerate(zip(offsets, sizes)):
        owner.append(b)
        end = offsets[i + 1] if i + 1 < len(offsets) else total
        if end - lo >= target and b < n_buckets - 1 and i + 1 < len(offsets):
            bounds.append((lo, end))
            lo, b = end, b + 1
    bounds.append((lo, total))
    return bounds, owner


class GradAllReduce:
    """Bucketed SUM all-reduce of `FusedAdam.flat_grad`, overlapped with backward.

        sync = GradAllReduce(fused_adam, n_buckets=3)
        loss.backward()            # hooks fire, buckets go out as they complete
        sync.finish()              # before optimizer.step()
    """

    def __init__(self, optimizer, n_buckets: int = 3, group=None, overlap: bool = True, backend: str = "nccl"):
        """backend "nccl": bucketed `dist.all_reduce`, overlapped with backward through post-accumulate hooks.
        backend "symm": the gradient arena is moved into NVLink symmetric memory (torch.distributed._symmetric_memory)
        and reduced by ONE peer-memory kernel after backward (multimem / two-shot): a captured NCCL collective costs a
        fixed ~0.1-0.3 ms per step on this path whatever its size, the peer-memory kernel a few tens of microseconds."""
        self.opt = optimizer
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.backend = backend if (self.world > 1 and optimizer.flat_grad.is_cuda) else "nccl"
        self.symm_op = None
        if self.backend == "symm":
            self._setup_symm()
            overlap = False
        params = optimizer._train
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]
        self.bounds, self.owner = plan_buckets(optimizer._offsets, sizes, optimizer.numel, n_buckets)
        self.expected = [self.owner.count(b) for b in range(len(self.bounds))]
        self.pending = list(self.expected)
        self.overlap = overlap and self.world > 1
        self.is_cuda = optimizer.flat_grad.is_cuda
        # PKA_COMM_PRIO=-1 puts the collectives on a high-priority stream (the priority is captured with the kernel nodes
        # of the step graph).  Measured at 2 GPUs (round 2, profiles/r02m_dp_exchange_n2.md): no difference, 1.010 vs
        # 1.014 ms per step -- the exposed cost is the last bucket's latency, not SM starvation -- so the default stays 0.
        prio = int(os.environ.get("PKA_COMM_PRIO", "0"))
        self.comm_stream = torch.cuda.Stream(priority=prio) if (self.is_cuda and self.overlap) else None
        self.launched = [False] * len(self.bounds)
        self.handles = []
        self.enabled = True
        if self.overlap:
            for i, p in enumerate(params):
                p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _setup_symm(self):
        import torch.distributed._symmetric_memory as symm_mem
        opt = self.opt
        pg = self.group if self.group is not None else dist.group.WORLD
        sym = symm_mem.empty(opt.numel, dtype=torch.float32, device=opt.flat_grad.device)
        sym.copy_(opt.flat_grad)
        self.symm_handle = symm_mem.rendezvous(sym, pg)
        self.group_name = pg.group_name
        for p, off in zip(opt._train, opt._offsets):          # the arena moved: re-point live .grad views
            if p.grad is not None and p.grad.data_ptr() == opt.flat_grad[off:off + 1].data_ptr():
                p.grad = sym[off:off + p.numel()].view(p.shape)
        opt.flat_grad = sym
        want = os.environ.get("PKA_SYMM_OP", "")
        ops_ns = torch.ops.symm_mem
        multicast = int(getattr(self.symm_handle, "multicast_ptr", 0) or 0) != 0      # NVSwitch multicast (NVLS) mapped?
        if want == "two_shot" or not multicast:
            self.symm_op = ops_ns.two_shot_all_reduce_
        else:
            self.symm_op = ops_ns.multimem_all_reduce_

    def _make_hook(self, i):
        def hook(_param):
            if not self.enabled:
                return
            b = self.owner[i]
            self.pending[b] -= 1
            if self.pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b):
        lo, hi = self.bounds[b]
        view = self.opt.flat_grad[lo:hi]
        self.launched[b] = True
        if self.opt.flat_grad.is_cuda:
            from . import ops
            ops.flush_deferred()                  # the bucket's gradients may still be split partials: sum them first
        if self.world == 1:
            return
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
        else:
            self.handles.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def paused(self):
        """Context manager: backward passes inside it exchange nothing (single-process reference runs on one rank, e.g.
        the DP-vs-single-process gradient check of bench.py)."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            self.enabled = False
            try:
                yield self
            finally:
                self.enabled = True
        return cm()

    def finish(self):
        """Send whatever has not gone out yet (no-overlap mode: everything), then join communication and compute."""
        if not self.enabled:
            return
        if self.opt.flat_grad.is_cuda:
            from . import ops
            ops.flush_deferred()
        if self.symm_op is not None:
            self.symm_op(self.opt.flat_grad, "sum", self.group_name)
            return
        for b in range(len(self.bounds)):
            if not self.launched[b]:
                self._launch(b)
        for h in self.handles:
            h.wait()
        self.handles = []
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.pending = list(self.expected)
        self.launched = [False] * len(self.bounds)

    __call__ = finish


def all_reduce_stats(totals: torch.Tensor, group=None) -> torch.Tensor:
    """Per-epoch SUM of (loss, n_correct, n_words) across ranks (the three accumulators of L/train.py:203-214)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(totals, op=dist.ReduceOp.SUM, group=group)
    return totals
