"""Training-loop side of the hot path with the reference's function names and signatures (L/train.py:58-214):
`cal_loss`, `get_performance`, `train_epoch`.  The step itself -- forward, summed CE (+accuracy), backward, Adam, LR
tick -- runs entirely as sm_100a kernels; `GraphedTrainStep` additionally captures it into one CUDA graph, which is
what makes the tiny TIMIT model (launch-bound) run at GPU speed.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops
from .utils import constants


def cal_loss(pred, goal, smoothing):
    """Summed cross-entropy over non-PAD targets; label smoothing eps=0.1 over the other V-1 classes when asked
    (L/train.py:72-90).  `pred` [N,V] logits, `goal` [N]."""
    loss, _ = ops.cross_entropy_sum(pred, goal.contiguous().view(-1), smoothing)
    return loss


def get_performance(crit, pred, goal, smoothing=True, num_class=None):
    """-> (loss_sum, n_correct) like L/train.py:58-69; both come out of ONE kernel pass over the logits."""
    pred2 = pred.contiguous().view(-1, pred.size(-1))
    loss, stats = ops.cross_entropy_sum(pred2, goal.contiguous().view(-1), smoothing)
    return loss, stats[0]


def initialize_batch_loader(read_feats_scp_file, read_text_file, read_vocab_file, batch_size, mode='drop', **loader_kw):
    """feats.scp + text + vocab -> BatchLoader over the utterances that have both features and a transcript
    (L/train.py:20-55): labels become [BOS] + vocabulary indices (UNK for unknown words) + [EOS].  `loader_kw` goes to
    `utils.BatchLoader.BatchLoader` (pad_to, bucket, seed, shard, ...); the defaults are the reference's: pre-loaded,
    silent, whole-set padding."""
    from .utils import instances_handler, kaldi_ark
    from .utils.BatchLoader import BatchLoader
    utterances = kaldi_ark.read_scp(read_feats_scp_file)
    print('[INFO] get {} utterances from {}.'.format(len(utterances), read_feats_scp_file))
    label_text = {}
    with open(read_text_file, encoding='utf-8') as f:
        for line in f:
            fields = line.split()
            if fields:
                label_text[fields[0]] = fields[1:]
    print('[INFO] get {} labels from {}.'.format(len(label_text), read_text_file))
    label = instances_handler.apply_vocab(instances_handler.add_control_words(label_text), read_vocab_file, 'word2idx')
    trainning_triples = [(key, rx, label[key]) for key, rx in utterances.items() if key in label]
    print('[INFO] match {} utterance-label pairs.'.format(len(trainning_triples)))
    loader_kw.setdefault('pre_load', True)
    loader_kw.setdefault('print_info', False)
    return BatchLoader(trainning_triples, batch_size, mode=mode, **loader_kw)


def _to_device(batch, device, non_blocking=True):
    """numpy batch tuple -> device tensors (the reference's FloatTensor/ByteTensor/LongTensor + .cuda(), L/train.py:151-161)."""
    def put(x, dtype):
        t = torch.as_tensor(np.ascontiguousarray(x), dtype=dtype) if not torch.is_tensor(x) else x.to(dtype)
        return t.to(device, non_blocking=non_blocking)
    return put(batch[1], torch.float32), put(batch[2], torch.uint8), put(batch[3], torch.int64), put(batch[4], torch.uint8)


class _Prefetcher:
    """Host->device staging a few batches ahead on a side stream: `take()` hands out the oldest staged batch,
    `release()` + `stage_next()` -- called right after the step's kernels have been launched and before any host sync --
    start the H2D copies of a later batch so they overlap the running steps (pinned host batches copy asynchronously).
    `depth + 1` sets of device staging buffers are reused while the batch shape stays the same, so the steady state
    performs no allocation.  The reference copies synchronously inside the step (L/train.py:151-161)."""

    _DTYPES = (torch.float32, torch.uint8, torch.int64, torch.uint8)
    _NP_DTYPES = (np.float32, np.uint8, np.int64, np.uint8)

    _CACHE = {}    # per device: copy stream, staging buffers, pinned read-back ring -- created once, reused by every epoch

    @classmethod
    def resources(cls, device, depth):
        key = (torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device(), depth)
        res = cls._CACHE.get(key)
        if res is None:
            res = dict(stream=torch.cuda.Stream(device=device), slots=[None] * (depth + 1),
                       host=[[None] * 4 for _ in range(depth + 1)], copied_ev=[None] * (depth + 1),
                       pinned=[torch.empty(3, dtype=torch.float32).pin_memory() for _ in range(depth + 1)])
            cls._CACHE[key] = res
        return res

    def __init__(self, batch_loader, device, depth=3):
        from collections import deque
        self.device = device
        res = self.resources(device, depth)
        self.copy_stream = res["stream"]
        self.copy_stream.wait_stream(torch.cuda.current_stream())     # buffers may still be read by the previous epoch
        self.it = iter(batch_loader)
        self.n_slots = depth + 1
        self.slots = res["slots"]              # device staging buffers
        self.host = res["host"]                # pinned host staging (flat, grow-only), used with loader.next_into
        self.copied_ev = res["copied_ev"]      # copy-stream event: this slot's H2D copy has left the host buffers
        self.next_into = getattr(self.it, "next_into", None)
        self.free_ev = [None] * self.n_slots   # compute-stream event: the consumer of this slot has finished reading it
        self.n = 0
        self.queue = deque()
        self.cur_slot = None
        for _ in range(depth):
            self.stage_next()

    def _pinned_alloc(self, slot):
        """Destination buffers for `loader.next_into`: views of this slot's grow-only pinned host memory, handed out once
        the slot's previous H2D copy has completed, so the loader writes each batch exactly once on the host."""
        def alloc(specs):
            if self.copied_ev[slot] is not None:
                self.copied_ev[slot].synchronize()
            flats = self.host[slot]
            views = []
            for k, (shape, dt) in enumerate(specs):
                tdt = self._DTYPES[k]
                if np.dtype(dt) != np.dtype(self._NP_DTYPES[k]):
                    raise TypeError("next_into: buffer %d must be %s" % (k, np.dtype(self._NP_DTYPES[k])))
                numel = int(np.prod(shape))
                if flats[k] is None or flats[k].numel() < numel:
                    flats[k] = torch.empty(max(numel, 1), dtype=tdt, pin_memory=True)
                views.append(flats[k][:numel].view(tuple(shape)))
            self._views = views
            return [v.numpy() for v in views]
        return alloc

    def stage_next(self):
        slot = self.n % self.n_slots
        try:
            if self.next_into is not None:
                self.next_into(self._pinned_alloc(slot))
                hosts = self._views
            else:
                batch = next(self.it)
                hosts = [x.to(dt) if torch.is_tensor(x) else torch.as_tensor(np.ascontiguousarray(x), dtype=dt)
                         for x, dt in zip(batch[1:5], self._DTYPES)]
        except StopIteration:
            return
        self.n += 1
        with torch.cuda.stream(self.copy_stream):
            if self.free_ev[slot] is not None:
                self.copy_stream.wait_event(self.free_ev[slot])
            bufs = self.slots[slot]
            if bufs is None or any(tuple(b.shape) != tuple(h.shape) for b, h in zip(bufs, hosts)):
                bufs = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in hosts]
                self.slots[slot] = bufs
            for b, h in zip(bufs, hosts):
                b.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.copied_ev[slot] = ev
        self.queue.append((tuple(bufs), ev, slot))

    def take(self):
        if not self.queue:
            return None
        tensors, ev, slot = self.queue.popleft()
        self.cur_slot = slot
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        for t in tensors:
            t.record_stream(cur)
        return tensors

    def release(self):
        """Everything that reads the batch handed out by the last take() has been enqueued on the current stream."""
        if self.cur_slot is not None:
            ev = torch.cuda.Event()
            ev.record()
            self.free_ev[self.cur_slot] = ev
            self.cur_slot = None


def train_epoch(model, batch_loader, crit, mode='train', optimizer=None, batch_eval=10, use_gpu=False, seq_error_prob=0,
                smoothing=False, graphed: Optional["GraphedTrainStep"] = None, sync_every_step: bool = False,
                grad_sync=None, stats_reduce=None):
    """One pass over `batch_loader` (L/train.py:127-214).  Returns (loss per word, token accuracy).

    Differences from the reference, all behaviour-preserving: the dead `use_seq_error` branches are not carried; the
    label-smoothing switch the reference hard-wires to False (L/train.py:193) is a keyword (default False); the three
    running totals stay on the device and are read back once per epoch (`sync_every_step=True` reads them back after
    every step like the reference does); host->device copies of batch i+1 overlap step i.  `grad_sync` (data parallel:
    `parallel.GradAllReduce(...).finish`) is called between backward and the optimiser step of the eager path; a
    `graphed` step carries its own; `stats_reduce` (`parallel.all_reduce_stats`) sums the epoch's three totals over the
    ranks before the ratios are taken (L/train.py:203-214 on the global batch).  This path has no CPU mode: the model must be on a CUDA device (`use_gpu` is
    accepted for signature compatibility)."""
    if mode == 'train':
        model.train()
        batch_loader.mode = 'drop'
    elif mode == 'eval':
        model.eval()
        batch_loader.mode = 'all'
        seen = 0
    else:
        raise ValueError('[ERROR] invalid epoch mode')
    device = next(model.parameters()).device
    if device.type != 'cuda':
        raise RuntimeError("train_epoch: the B200 path has no CPU mode; move the model to a CUDA device first")
    totals = torch.zeros(3, device=device, dtype=torch.float64)          # loss, n_correct, n_words
    host_totals = [0.0, 0.0, 0.0]     # sync_every_step: the reference reads loss/accuracy back on every step (L/train.py:203-207)

    depth = 3                         # steps the host may run ahead of the GPU (absorbs host scheduling hiccups)
    feed = _Prefetcher(batch_loader, device, depth)
    # sync_every_step: each step's [loss, n_correct, n_words] is copied to pinned host memory right behind the step and
    # consumed a few steps later, so the host never drains the GPU queue between steps
    pinned = _Prefetcher.resources(device, depth)["pinned"] if sync_every_step else None
    from collections import deque
    pending, n_step = deque(), 0

    def read_back(vec3):
        nonlocal n_step
        buf = pinned[n_step % (depth + 1)]
        n_step += 1
        buf.copy_(vec3, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        pending.append((buf, ev))
        while len(pending) > depth:
            consume()

    def consume():
        buf, ev = pending.popleft()
        ev.synchronize()
        for k, v in enumerate(buf.tolist()):
            host_totals[k] += v

    def flush():
        while pending:
            consume()

    while True:
        cur = feed.take()
        if cur is None:
            break
        src_seq, src_pad_mask, tgt_seq, tgt_pad_mask = cur
        if graphed is not None and mode == 'train':
            out = graphed.step(src_seq, src_pad_mask, tgt_seq, tgt_pad_mask)
            feed.release()
            feed.stage_next()
            if sync_every_step:
                read_back(out)
            else:
                totals += out.double()
            continue
        tgt_in, goal, tgt_in_mask = ops.split_targets(tgt_seq, tgt_pad_mask)     # L/train.py:163-165, one launch
        if mode == 'train':
            optimizer.zero_grad()
            pred = model(src_seq, src_pad_mask, tgt_in, tgt_in_mask)
        else:
            with torch.no_grad():
                pred = model(src_seq, src_pad_mask, tgt_in, tgt_in_mask)
        loss, stats = ops.cross_entropy_sum(pred.view(-1, pred.size(-1)), goal.view(-1), smoothing)
        if mode == 'train':
            loss.backward()
            if grad_sync is not None:
                grad_sync()
            optimizer.step()
            optimizer.update_learning_rate()
        feed.release()
        feed.stage_next()
        if sync_every_step:
            read_back(torch.cat([loss.detach().reshape(1), stats.reshape(2)]))
        else:
            totals[0] += loss.detach().double()
            totals[1:] += stats.double()
        if mode == 'eval':
            seen += 1
            if seen == batch_eval:
                break
    flush()
    if stats_reduce is not None:
        totals = stats_reduce(totals + torch.tensor(host_totals, dtype=torch.float64, device=device))
        host_totals = [0.0, 0.0, 0.0]
    loss_sum, n_correct, n_words = [a + b for a, b in zip(totals.tolist(), host_totals)]
    return loss_sum / int(n_words), n_correct / int(n_words)


class GraphedTrainStep:
    """Whole training step (zero_grad, forward, loss, backward, [gradient all-reduce], Adam, LR tick, dropout tick)
    captured into CUDA graphs -- one per batch shape -- and replayed per batch.

    The reference's pre-loading BatchLoader pads the whole data set to one length (U/BatchLoader.py:33-36): one shape,
    one graph.  With per-batch padding to a few bucket lengths (`utils.BatchLoader(pad_to="bucket", bucket=...)`) every
    bucket shape gets its own graph, captured the first time the shape is seen (or up front with `capture`); the graphs
    share one memory pool (they never run concurrently and no tensor created inside one outlives its replay).

    Inputs are copied into the shape's static device buffers; `step` returns a device tensor
    [loss_sum, n_correct, n_words]."""

    def __init__(self, model, optimizer, example_batch=None, smoothing=False, grad_sync=None, warmup=3):
        self.model, self.optimizer, self.smoothing, self.grad_sync = model, optimizer, smoothing, grad_sync
        self.device = next(model.parameters()).device
        self.out = torch.zeros(3, device=self.device, dtype=torch.float32)
        self._one = torch.ones((), device=self.device, dtype=torch.float32)
        self.warmup = warmup
        self.shapes = {}                                  # (B, T, F, L+1) -> dict(src, smask, tgt, tmask, graph)
        self.pool = None
        self.cur = None
        if example_batch is not None:
            self.capture(example_batch)

    # -- the most recently used shape, under the names the single-shape version exposed
    @property
    def graph(self):
        return self.cur["graph"]

    @property
    def src(self):
        return self.cur["src"]

    @property
    def tgt(self):
        return self.cur["tgt"]

    @staticmethod
    def _key(src, tgt):
        return tuple(src.shape) + (int(tgt.shape[1]),)

    def capture(self, batch):
        """Make sure a graph exists for this batch's shape (batch tuple of numpy arrays or device tensors)."""
        if len(batch) == 5:
            batch = batch[1:]
        key = self._key(batch[0], batch[2])
        ent = self.shapes.get(key)
        if ent is None:
            ent = self.shapes[key] = self._capture(batch)
        self.cur = ent
        return ent

    def _capture(self, batch):
        model, device = self.model, self.device
        src, smask, tgt, tmask = _to_device((None,) + tuple(batch), device, non_blocking=False)
        ent = dict(src=src.clone(), smask=smask.clone(), tgt=tgt.clone(), tmask=tmask.clone(), graph=None)
        was_training = model.training
        model.train()
        snap = self._snapshot()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup if not self.shapes else 1):     # warm allocator pools & lazy state outside the capture
                self._body(ent)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if self.pool is None:
            self.pool = torch.cuda.graph_pool_handle()
        ent["graph"] = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ent["graph"], pool=self.pool):
            self._body(ent)
        self._restore(snap)                               # warm-up steps must not count as training
        model.train(was_training)
        return ent

    def _snapshot(self):
        inner = getattr(self.optimizer, "optimizer", self.optimizer)
        if not hasattr(inner, "flat_param"):
            raise RuntimeError("GraphedTrainStep needs a FusedAdam (flat parameter arena) inside the optimizer")
        rng = self.model.dropout_state
        dev = self.device
        return dict(tensors=[(t, t.clone()) for t in (inner.flat_param, inner.exp_avg, inner.exp_avg_sq, inner.dev_state,
                                                      inner.dev_lr, rng.step_tensor(dev), self.out)],
                    n=getattr(self.optimizer, "n_current_steps", None), lr=[g["lr"] for g in inner.param_groups])

    def _restore(self, snap):
        inner = getattr(self.optimizer, "optimizer", self.optimizer)
        for t, saved in snap["tensors"]:
            t.copy_(saved)
        if inner.flat_shadow is not None:
            inner.flat_shadow.copy_(inner.flat_param)
        if snap["n"] is not None:
            self.optimizer.n_current_steps = snap["n"]
        for g, lr in zip(inner.param_groups, snap["lr"]):
            g["lr"] = lr
        torch.cuda.synchronize()

    def _body(self, ent):
        tgt_in, goal, tmask_in = ops.split_targets(ent["tgt"], ent["tmask"])
        self.optimizer.zero_grad()
        pred = self.model(ent["src"], ent["smask"], tgt_in, tmask_in)
        # the loss kernel writes {loss, n_correct, n_words} straight into the step's result buffer
        loss, _ = ops.cross_entropy_sum(pred.view(-1, pred.size(-1)), goal.view(-1), self.smoothing, out3=self.out)
        loss.backward(self._one)                  # preallocated d(loss) = 1: no fill kernel per step
        if self.grad_sync is not None:
            self.grad_sync()
        self.optimizer.step()
        self.optimizer.update_learning_rate()

    def load(self, src, smask, tgt, tmask):
        ent = self.shapes.get(self._key(src, tgt))
        if ent is None:
            ent = self.capture((src, smask, tgt, tmask))
        self.cur = ent
        ent["src"].copy_(src, non_blocking=True)
        ent["smask"].copy_(smask, non_blocking=True)
        ent["tgt"].copy_(tgt, non_blocking=True)
        ent["tmask"].copy_(tmask, non_blocking=True)
        return ent

    def step(self, src, smask, tgt, tmask):
        ent = self.load(src, smask, tgt, tmask)
        ent["graph"].replay()
        self._host_tick()
        return self.out

    def _host_tick(self):
        """Keep the host-side mirror of the LR schedule in step with the device counters the graph advances (a
        checkpoint written after graphed training must carry the right schedule step; the arithmetic is
        ScheduledOptim.update_learning_rate's, no kernel)."""
        opt = self.optimizer
        if hasattr(opt, "n_current_steps") and hasattr(opt, "soft_coefficient"):
            opt.n_current_steps += 1
            new_lr = (opt.start_lr * opt.soft_coefficient) / (opt.n_current_steps + opt.soft_coefficient)
            for group in opt.optimizer.param_groups:
                group["lr"] = new_lr

    def release(self):
        """Drop the captured graphs (before tearing down a process group whose collectives they contain)."""
        for ent in self.shapes.values():
            ent["graph"] = None
        self.shapes.clear()
        self.cur = None


# ------------------------------------------------------------------------------------------------ epoch driver
def get_criterion(vocab_size):
    """Kept for call-site compatibility (L/train.py:326-330): the loss kernel ignores PAD targets itself, `crit` is unused."""
    weight = torch.ones(vocab_size)
    weight[constants.PAD] = 0
    return torch.nn.CrossEntropyLoss(weight, reduction='sum')


def train(model, train_data, dev_data, test_data, crit, optimizer, opt, model_options, graphed=None, grad_sync=None,
          writer=True, stats_reduce=None):
    """Epoch loop of L/train.py:217-272: train, evaluate 10 training batches / dev / test, checkpoint every
    `opt.save_interval` epochs and every epoch of the last interval, finally write the best-on-dev model.
    -> (best_accu, best_epoch).

    Differences: checkpoints are state-dict files with optimiser / schedule / dropout state (checkpoint.py), so a run
    can resume at an epoch boundary with the same weights, optimiser / schedule / dropout state, epoch-shuffle RNG and
    best-so-far accuracy (the best *weights* of an earlier epoch are re-read from that epoch's checkpoint when it still
    exists); the best model is a *snapshot* taken at its epoch (the reference keeps a reference to the
    live module, L/train.py:243-245, and therefore saves the last epoch's weights under the best epoch's name).
    Data parallel: every rank calls this with its shard of `train_data` and `grad_sync`; `writer` is True on one rank.
    Like the reference, every evaluation stops after `batch_eval` = 10 batches (L/train.py:127,209-212)."""
    import time
    from . import checkpoint as _ckpt
    start_all, best_epoch, best_accu, best_state = time.time(), 0, 0.0, None
    first = int(getattr(opt, 'start_epoch', 1))
    resumed = getattr(opt, 'resume_extra', None) or {}
    if resumed:                                   # -resume: the epoch shuffles and the best-so-far continue where they stopped
        best_accu, best_epoch = float(resumed.get('best_accu', 0.0)), int(resumed.get('best_epoch', 0))
        best_state = resumed.get('best_state')
        rng_state = resumed.get('loader_rng')
        if rng_state is not None and hasattr(train_data, 'set_rng_state'):
            train_data.set_rng_state(rng_state)
    for epoch in range(first, opt.epoch + 1):
        print('[INFO] trainning epoch {}.'.format(epoch))
        start = time.time()
        _, train_accu = train_epoch(model, train_data, crit, mode='train', optimizer=optimizer, use_gpu=True,
                                    seq_error_prob=getattr(opt, 'seq_error_prob', 0), graphed=graphed, grad_sync=grad_sync,
                                    stats_reduce=stats_reduce)
        print('[INFO]-----(Training)----- accuracy: {:3.2f} %, elapse: {:3.2f} min'
              .format(100 * train_accu, (time.time() - start) / 60))
        start = time.time()
        eval_batch_num = 10
        _, accu = train_epoch(model, train_data, crit, mode='eval', batch_eval=eval_batch_num, use_gpu=True)
        print('[INFO]-----(evaluating train set for {} batch)----- accuracy: {:3.2f} %, elapse: {:3.2f} min'
              .format(eval_batch_num, 100 * accu, (time.time() - start) / 60))
        start = time.time()
        _, valid_accu = train_epoch(model, dev_data, crit, mode='eval', use_gpu=True)
        print('[INFO]-----(evaluating dev set)----- accuracy: {:3.2f} %, elapse: {:3.2f} min'
              .format(100 * valid_accu, (time.time() - start) / 60))
        if valid_accu > best_accu:
            best_accu, best_epoch = valid_accu, epoch
            best_state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        start = time.time()
        _, test_accu = train_epoch(model, test_data, crit, mode='eval', use_gpu=True)
        print('[INFO]-----(evaluating test set)----- accuracy: {:3.2f} %, elapse: {:3.2f} min'
              .format(100 * test_accu, (time.time() - start) / 60))
        due = epoch % opt.save_interval == 0 or opt.epoch - epoch < opt.save_interval
        inner = getattr(optimizer, 'optimizer', optimizer)
        if due and getattr(inner, '_peer', None) is not None:
            inner.sync_moments()                      # peer mode shards the Adam moments: gather them (all ranks)
        if writer and due:
            model_name = opt.save_model_dir + '/epoch.{}.torch'.format(epoch)
            extra = dict(best_accu=float(best_accu), best_epoch=int(best_epoch))
            if hasattr(train_data, 'get_rng_state'):
                extra['loader_rng'] = train_data.get_rng_state()
            _ckpt.save_checkpoint(model_name, model, model_options, epoch, train_options=opt, optimizer=optimizer, extra=extra)
            print('[INFO] checkpoint of epoch {} is saved to {}'.format(epoch, model_name))
    print('[INFO] trainning finish.\n\ttime consume: {:3.2f} minute\n\tbest valid accuracy: {:3.2f} %, on epoch {}'
          .format((time.time() - start_all) / 60, 100 * best_accu, best_epoch))
    if best_state is not None and writer:
        model_name = opt.save_model_dir + '/best.epoch{}.accu{:3.2f}.torch'.format(best_epoch, 100 * best_accu)
        _ckpt.save_state(model_name, best_state, model, model_options, best_epoch, train_options=opt)
        print('[INFO] best model is saved to {}'.format(model_name))
    return best_accu, best_epoch


def combine(opt, epoch, crit, data, num_model=20, device='cuda'):
    """Model averaging of L/train.py:284-322: running mean over the checkpoints of epochs `epoch`, `epoch-1`, ...
    (`num_model` of them, missing files end the walk), evaluated on `data` after every addition; the best average is
    written to `combined.accuXX.XX.torch` under `opt.save_model_dir`.  -> best accuracy."""
    import os
    files = []
    for i in range(epoch, epoch - num_model, -1):
        name = opt.save_model_dir + '/epoch.{}.torch'.format(i)
        if not os.path.exists(name):
            break
        files.append(name)
    if not files:
        raise ValueError('[ERROR] no checkpoint epoch.{}.torch under {}'.format(epoch, opt.save_model_dir))
    return combine_files(files, crit, data, opt.save_model_dir, train_options=opt, device=device)


def combine_files(files, crit, data, save_model_dir, train_options=None, device='cuda'):
    """The averaging loop shared by `combine` and the stand-alone L/combine.py:33-108: running mean of the model files
    in the given order, evaluation on `data` after every addition, best average saved.  -> best accuracy.

    The running mean lives on the device in the reference's arithmetic (checkpoint.running_average); the best average is
    a snapshot (the reference's `best_model` aliases the live module, so it saves the last average whatever its score)."""
    import math
    import time
    from . import checkpoint as _ckpt
    print('[PROCEDURE] combining model with model averaging...')
    first = _ckpt.load_checkpoint(files[0], device=device)
    model, model_options = first['model'], first['model_options']
    print('[INFO] model loaded')

    def states():
        yield {k: v.to(device) for k, v in first['state_dict'].items()}
        for name in files[1:]:
            yield {k: v.to(device) for k, v in _ckpt.read_checkpoint(name)['state_dict'].items()}

    best_accu, best_state, best_n = -1.0, None, 0
    for n, avg in _ckpt.running_average(states()):
        print('[INFO] averaging {} models'.format(n))
        model.load_state_dict(avg)
        start = time.time()
        test_loss, test_accu = train_epoch(model, data, crit, mode='eval', use_gpu=True)
        print('[INFO]-----(evaluating combining set)----- ppl: {:7.3f}, accuracy: {:3.2f} %, elapse: {:3.2f} min'
              .format(math.exp(min(test_loss, 100)), 100 * test_accu, (time.time() - start) / 60))
        if test_accu > best_accu:
            best_accu, best_n = test_accu, n
            best_state = {k: v.detach().cpu().clone() for k, v in avg.items()}
    print('[INFO] best combined model with accuracy: {:3.2f} %'.format(100 * best_accu))
    model_name = save_model_dir + '/combined.accu{:3.2f}.torch'.format(100 * best_accu)
    _ckpt.save_state(model_name, best_state, model, model_options, first['epoch'], train_options=train_options,
                     extra=dict(averaged_models=best_n, averaged_from=list(files[:best_n])))
    return best_accu
