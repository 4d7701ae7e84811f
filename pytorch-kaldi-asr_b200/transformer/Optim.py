"""Optimiser side of the training step: the reference's `ScheduledOptim` wrapper (T/Optim.py:4-27) plus `FusedAdam`,
which replaces torch.optim.Adam's per-tensor update over 72 tensors by ONE kernel over a flat parameter arena.

Arena layout (HBM): all trainable parameters live back to back in one fp32 buffer (`flat_param`), their gradients in a
second one (`flat_grad`, `p.grad` are views into it), Adam moments in two more.  One launch updates everything; under
data parallelism the same flat gradient buffer is what NCCL all-reduces (parallel.py), in a few large buckets.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib as L


class FusedAdam(torch.optim.Optimizer):
    """Adam(betas, eps, no weight decay, no amsgrad) with torch.optim.Adam's arithmetic (bias corrections in double,
    denom = sqrt(v)/sqrt(bc2) + eps, step = lr/bc1), fused over the arena.  Parameters must already be on the GPU."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, bf16_shadow=False):
        params = [p for p in params]
        # weight_decay / amsgrad are fixed at torch.optim.Adam's defaults; they are listed so that this optimizer's
        # state_dict() loads into a stock Adam (which reads both keys from every parameter group)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False))
        self._train = [p for g in self.param_groups for p in g["params"] if p.requires_grad]
        if not self._train:
            raise ValueError("FusedAdam: no trainable parameters")
        dev = self._train[0].device
        L.require_cuda(*self._train)
        for p in self._train:
            if p.dtype != torch.float32:
                raise RuntimeError("FusedAdam: fp32 master parameters expected")
        # 16-byte align every tensor inside the arena so kernels may use 128-bit accesses on the views
        self._offsets, total = [], 0
        for p in self._train:
            self._offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.numel = total
        self.flat_param = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_shadow = torch.zeros(total, device=dev, dtype=torch.bfloat16) if bf16_shadow else None
        with torch.no_grad():
            for p, off in zip(self._train, self._offsets):
                view = self.flat_param[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_grad[off:off + p.numel()].view(p.shape)
            if self.flat_shadow is not None:
                self.flat_shadow.copy_(self.flat_param)
        from .. import ops as _ops
        _ops.register_grad_slots(self)            # backward kernels write gradients straight into the arena slots
        self.dev_state = torch.zeros(2, device=dev, dtype=torch.int64)        # {adam_t, n_current_steps}
        self.dev_lr = torch.full((1,), float(lr), device=dev, dtype=torch.float32)
        self.use_device_lr = False                                            # switched on by ScheduledOptim
        self._peer = None                                                     # set by enable_peer_step()
        self._done = torch.zeros(1, device=dev, dtype=torch.int32)            # CTA-completion counter of the fused step

    # ---- data parallel: gradient reduce-scatter + Adam + parameter all-gather in one kernel over peer memory ----
    def enable_peer_step(self, group=None, max_ctas: int = 148, multicast=None):
        """Move the parameter and gradient arenas into NVLink symmetric memory (the same allocation mapped on every
        rank of `group`) and make `step()` one `pka_dp_adam_step` launch per rank: the gradients of all ranks are summed
        while being read (through the switch when NVLS multicast is available), this rank updates its 1/W shard and
        writes the new parameters to every rank.  No separate all-reduce: do not combine with `GradAllReduce`.
        Call on every rank, after construction and before `GraphedTrainStep` captures the step.  `multicast=False`
        (or PKA_DP_MULTICAST=0) forces the peer-pointer path, whose summation order is fixed (bit-reproducible).
        Adam moments then live sharded: call `sync_moments()` on every rank before `state_dict()`."""
        import os
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from .. import ops as _ops
        if self.flat_shadow is not None:
            raise RuntimeError("enable_peer_step: the bf16 shadow arena is not supported in peer mode")
        pg = group if group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(pg), dist.get_rank(pg)
        dev = self.flat_param.device
        quantum = 4 * world
        n = (self.numel + quantum - 1) // quantum * quantum

        def regrow(old, symmetric):
            new = symm_mem.empty(n, dtype=torch.float32, device=dev) if symmetric else \
                torch.empty(n, dtype=torch.float32, device=dev)
            new.zero_()
            new[:old.numel()].copy_(old)
            return new

        with torch.no_grad():
            new_param, new_grad = regrow(self.flat_param, True), regrow(self.flat_grad, True)
            for p, off in zip(self._train, self._offsets):
                old_slot = self.flat_grad[off:off + p.numel()]
                keeps_view = p.grad is not None and p.grad.data_ptr() == old_slot.data_ptr()
                p.data = new_param[off:off + p.numel()].view(p.shape)
                if keeps_view:
                    p.grad = new_grad[off:off + p.numel()].view(p.shape)
            self.exp_avg, self.exp_avg_sq = regrow(self.exp_avg, False), regrow(self.exp_avg_sq, False)
            self.flat_param, self.flat_grad, self.numel = new_param, new_grad, n
        _ops.register_grad_slots(self)                                        # the parameters moved
        h_param, h_grad = symm_mem.rendezvous(new_param, pg), symm_mem.rendezvous(new_grad, pg)
        grid = int(L.lib().pka_dp_adam_grid(C.c_int64(n), world, max_ctas))
        flags = symm_mem.empty(2 * grid * world, dtype=torch.int32, device=dev)
        flags.zero_()
        h_flags = symm_mem.rendezvous(flags, pg)
        if multicast is None:
            multicast = os.environ.get("PKA_DP_MULTICAST", "1") != "0"

        def addresses(handle, tensor):
            """Peer addresses and multicast address of `tensor` (the handle describes the allocation it lives in)."""
            off = tensor.data_ptr() - int(handle.buffer_ptrs[rank])    # position inside the symmetric allocation
            if off < 0 or off + tensor.numel() * tensor.element_size() > int(handle.buffer_size):
                raise RuntimeError("enable_peer_step: symmetric-memory handle does not describe this tensor "
                                   "(local address %#x, allocation at %#x + %d bytes)"
                                   % (tensor.data_ptr(), int(handle.buffer_ptrs[rank]), int(handle.buffer_size)))
            peers = [int(a) + off for a in handle.buffer_ptrs]
            mc = int(getattr(handle, "multicast_ptr", 0) or 0)
            return peers, (mc + off if (mc and multicast) else 0)

        param_peers, mc_param = addresses(h_param, new_param)
        grad_peers, mc_grad = addresses(h_grad, new_grad)
        flag_peers, _ = addresses(h_flags, flags)
        if not (mc_param and mc_grad):
            mc_param = mc_grad = 0
        self._peer = dict(rank=rank, world=world, max_ctas=int(max_ctas), group=pg, handles=(h_param, h_grad, h_flags),
                          flags=flags, param_ptrs=(C.c_uint64 * world)(*param_peers),
                          grad_ptrs=(C.c_uint64 * world)(*grad_peers), param_mc=mc_param, grad_mc=mc_grad,
                          flag_ptrs_dev=torch.tensor(flag_peers, dtype=torch.int64, device=dev))
        torch.cuda.synchronize(dev)
        dist.barrier(pg)                                                      # every rank's flags are zero before any launch
        return self

    @torch.no_grad()
    def sync_moments(self):
        """Peer mode keeps each rank's Adam moments valid only on its shard; gather them (collective: all ranks)."""
        if self._peer is None:
            return
        import torch.distributed as dist
        pr = self._peer
        for arena in (self.exp_avg, self.exp_avg_sq):
            shard = arena.view(pr["world"], -1)[pr["rank"]].clone()
            dist.all_gather_into_tensor(arena, shard, group=pr["group"])

    # ---- state interchange: torch.optim.Adam's own layout, so checkpoints move between FusedAdam and stock Adam ----
    def _param_index(self):
        return {id(p): i for i, p in enumerate(q for g in self.param_groups for q in g["params"])}

    def state_dict(self):
        """{'state': {param index: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]} as torch.optim.Adam writes
        it (per-parameter copies of the arena moments; empty before the first step), plus 'fused': the device-side
        counters {adam_t, n_current_steps, lr}.  Synchronises the stream (checkpoint time only)."""
        base = super().state_dict()
        adam_t, n_sched = (int(v) for v in self.dev_state.tolist())
        state = {}
        if adam_t > 0:
            index = self._param_index()
            for p, off in zip(self._train, self._offsets):
                sl = slice(off, off + p.numel())
                state[index[id(p)]] = dict(step=torch.tensor(float(adam_t)),
                                           exp_avg=self.exp_avg[sl].view(p.shape).clone(),
                                           exp_avg_sq=self.exp_avg_sq[sl].view(p.shape).clone())
        base["state"] = state
        base["fused"] = dict(adam_t=adam_t, n_current_steps=n_sched, lr=float(self.dev_lr.item()))
        return base

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        if len(groups) != len(self.param_groups) or \
                any(len(a["params"]) != len(b["params"]) for a, b in zip(groups, self.param_groups)):
            raise ValueError("FusedAdam.load_state_dict: parameter groups do not match this optimizer")
        for mine, theirs in zip(self.param_groups, groups):
            for key in ("lr", "betas", "eps"):
                if key in theirs:
                    mine[key] = tuple(theirs[key]) if key == "betas" else theirs[key]
        index = self._param_index()
        state = {int(k): v for k, v in state_dict.get("state", {}).items()}
        steps = set()
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for p, off in zip(self._train, self._offsets):
            entry = state.get(index[id(p)])
            if entry is None:
                continue
            if tuple(entry["exp_avg"].shape) != tuple(p.shape):
                raise ValueError("FusedAdam.load_state_dict: moment shape %s does not match parameter shape %s"
                                 % (tuple(entry["exp_avg"].shape), tuple(p.shape)))
            sl = slice(off, off + p.numel())
            self.exp_avg[sl].view(p.shape).copy_(entry["exp_avg"])
            self.exp_avg_sq[sl].view(p.shape).copy_(entry["exp_avg_sq"])
            steps.add(int(float(entry["step"])))
        if len(steps) > 1:
            raise ValueError("FusedAdam.load_state_dict: parameters with different step counts %s (one shared "
                             "bias-correction step is kept for the whole arena)" % sorted(steps))
        fused = state_dict.get("fused", {})
        adam_t = int(fused.get("adam_t", steps.pop() if steps else 0))
        self.dev_state.copy_(torch.tensor([adam_t, int(fused.get("n_current_steps", 0))], dtype=torch.int64))
        self.dev_lr.fill_(float(fused.get("lr", self.param_groups[0]["lr"])))

    @torch.no_grad()
    def set_schedule_step(self, n_current_steps, lr):
        """Device-side copy of ScheduledOptim's counter and current learning rate (checkpoint restore)."""
        self.dev_state[1] = int(n_current_steps)
        self.dev_lr.fill_(float(lr))

    def grad_view(self, p):
        i = next(k for k, q in enumerate(self._train) if q is p)
        off = self._offsets[i]
        return self.flat_grad[off:off + p.numel()].view(p.shape)

    def shadow_view(self, p):
        i = next(k for k, q in enumerate(self._train) if q is p)
        off = self._offsets[i]
        return self.flat_shadow[off:off + p.numel()].view(p.shape)

    def _adopt_grads(self):
        """Make sure every p.grad *is* its arena view (a foreign .grad tensor is copied in; None counts as zero)."""
        for p, off in zip(self._train, self._offsets):
            view = self.flat_grad[off:off + p.numel()].view(p.shape)
            if p.grad is None:
                view.zero_()
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
            else:
                continue
            p.grad = view

    def zero_grad(self, set_to_none: bool = True):
        """Default: drop the .grad references (no kernel at all).  The next backward writes every gradient straight into
        its arena slot and autograd adopts those views; `step()` zero-fills the slot of any parameter that got none.
        set_to_none=False keeps torch's classic behaviour (one memset of the arena, gradients accumulate into it)."""
        from .. import ops as _ops
        _ops._SLOT_CLAIMS.clear()                 # a new backward pass may claim every arena slot again
        if set_to_none:
            for p in self._train:
                p.grad = None
            return
        self._adopt_grads()
        self.flat_grad.zero_()

    @torch.no_grad()
    def step(self, closure=None, lr_tick=None):
        """`lr_tick` = (start_lr, soft_coefficient): also advance the device-side LR schedule in the same launch (what
        ScheduledOptim.update_learning_rate would otherwise do with a kernel of its own); returns True if it did."""
        assert closure is None
        from .. import ops as _ops
        _ops.flush_deferred()                     # normally a no-op: the backward pass has flushed its own reductions
        self._adopt_grads()
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        lr_dev = L.ptr(self.dev_lr) if self.use_device_lr else C.c_void_p(0)
        if self._peer is not None:
            pr = self._peer
            L.check(L.lib().pka_dp_adam_step(pr["param_ptrs"], pr["grad_ptrs"], C.c_uint64(pr["param_mc"]),
                                             C.c_uint64(pr["grad_mc"]), L.ptr(pr["flag_ptrs_dev"]), pr["rank"], pr["world"],
                                             pr["max_ctas"], L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq),
                                             C.c_int64(self.numel), lr_dev, C.c_float(g["lr"]), L.ptr(self.dev_state),
                                             C.c_float(b1), C.c_float(b2), C.c_float(g["eps"]), L.stream_ptr()),
                    "dp_adam_step")
            return
        tick = lr_tick is not None and self.use_device_lr
        L.check(L.lib().pka_adam_step_fused(L.ptr(self.flat_param), L.ptr(self.flat_grad), L.ptr(self.exp_avg),
                                            L.ptr(self.exp_avg_sq), C.c_int64(self.numel), lr_dev, C.c_float(g["lr"]),
                                            L.ptr(self.dev_state), C.c_float(b1), C.c_float(b2), C.c_float(g["eps"]),
                                            L.ptr(self.flat_shadow), L.ptr(self._done), int(tick),
                                            C.c_float(lr_tick[0] if tick else 0.0), C.c_float(lr_tick[1] if tick else 0.0),
                                            L.stream_ptr()), "adam_step_fused")
        return tick


class ScheduledOptim(object):
    """lr_n = start_lr * c / (n + c), installed AFTER step n (so step 1 runs at the inner optimiser's constructor lr),
    exactly as T/Optim.py:21-27.  With a FusedAdam inside, the schedule also advances on the device (one tiny kernel),
    which keeps a captured CUDA graph of the whole step valid across replays."""

    def __init__(self, optimizer, start_lr=0.001, soft_coefficient=500):
        self.optimizer = optimizer
        self.start_lr = start_lr
        self.soft_coefficient = soft_coefficient
        self.n_current_steps = 0
        self._fused = isinstance(optimizer, FusedAdam)
        if self._fused:
            optimizer.use_device_lr = True

    def step(self):
        if self._fused:
            # the device-side schedule tick rides in the Adam launch; update_learning_rate() then only does the host math
            self._ticked = bool(self.optimizer.step(lr_tick=(self.start_lr, self.soft_coefficient)))
        else:
            self.optimizer.step()

    def zero_grad(self):
        self.optimizer.zero_grad()

    def update_learning_rate(self):
        """LR rule of T/Optim.py:21-27.  With a FusedAdam the device-side copy is normally advanced by step() itself;
        calling this without a preceding step() (or after a peer-mode step) still ticks the device."""
        self.n_current_steps += 1
        new_lr = (self.start_lr * self.soft_coefficient) / (self.n_current_steps + self.soft_coefficient)
        for group in self.optimizer.param_groups:
            group["lr"] = new_lr
        if self._fused and not getattr(self, "_ticked", False):
            o = self.optimizer
            L.check(L.lib().pka_lr_tick(L.ptr(o.dev_lr), L.ptr(o.dev_state), C.c_float(self.start_lr),
                                        C.c_float(self.soft_coefficient), L.stream_ptr()), "lr_tick")
        self._ticked = False
