#!/usr/bin/env python
"""Benchmark of the acoustic-model hot path (BASELINE.json): TIMIT example model, train step (forward + summed CE +
backward + Adam + LR schedule) on synthetic TIMIT-shaped batches of 32 padded utterances per GPU.

    python bench.py --gpus N --steps K --warmup W            # B200 path (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU (oracle port)

Prints ONE JSON line (rank 0).  `value` = real (unpadded) source frames per second over all ranks with inputs resident
in HBM; `e2e` = the same metric through the public API (`train_epoch`) from pinned host batches, H2D copies and the
per-step D2H read of the loss inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "train_frames_per_sec"
UNIT = "frames/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_tflops_sustained=p["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def wait_first(self, timeout=5.0):
        """nvidia-smi needs ~0.5 s before its first line; the timed region must not be over before it starts sampling."""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)
        self.rows.clear()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.kill()                   # make sure the poller is gone before the next timed region starts
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- CPU arm (oracle)
def cpu_train_baseline(n_utt: int, steps: int, warmup: int, threads: int):
    """The reference algorithm (oracle port: torch-CPU fp32, dropout 0.35 through torch's generator like nn.Dropout)
    timed on the host cores: forward + loss + backward + Adam + LR tick on `n_utt` utterances per step."""
    from oracle import acoustic_model as am
    from oracle import train_step as otrain
    from pytorch_kaldi_asr_b200.utils import synthetic
    torch.set_num_threads(threads)
    cfg = am.example_config()
    sd = am.init_state_dict(cfg, synthetic.lda_matrix(), seed=0)
    pool = synthetic.batches(2, n_utt, seed=1234)
    opt = otrain.AdamSchedule({k: sd[k] for k in am.trainable_keys(sd)}, 1e-3, 25000)

    def one(batch):
        _, loss, _, _, grads = otrain.loss_and_grads(sd, cfg, batch[1:], False, am.DropoutPlan("rng"))
        opt.step(grads)
        opt.update_learning_rate()
        return float(loss)

    for i in range(warmup):
        one(pool[i % 2])
    frames, t0 = 0, time.perf_counter()
    for i in range(steps):
        one(pool[i % 2])
        frames += synthetic.real_frames(pool[i % 2])
    dt = time.perf_counter() - t0
    return frames / dt, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_utt = 32
    value, ms = cpu_train_baseline(n_utt, args.steps, min(args.warmup, 2), threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 2), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "timit_example_model_train_step_B32 (configs[1])", "global_batch": n_utt,
                   "batch_per_step": n_utt},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d steps x %d utterances, oracle port of L/train.py:145-207, dropout 0.35" % (args.steps, n_utt)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------- B200 arm
def ensure_library():
    import importlib.util
    spec = importlib.util.spec_from_file_location("pka_build", os.path.join(ROOT, "pytorch-kaldi-asr_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not os.path.exists(mod.LIB):
        mod.build_library()


def time_kernel(fn, iters=20, warm=3, replays=5):
    """Average device time of one launch: `iters` launches are captured into a CUDA graph (so the Python/ctypes
    launch overhead of this harness is not what is measured) and the replays are bracketed by CUDA events on the
    launching stream."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(replays):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (iters * replays) * 1e-3          # seconds per launch


def roofline_probe(model, B, T, pk, mode):
    """Dominant kernel of the step = the TDNN GEMM family (6 forward + 6 dgrad + 6 wgrad launches, ~85 % of the FLOPs).
    Timed alone with CUDA events on the launching stream at the step's real shape [B*T, 768] x [768, 256];
    algorithmic FLOPs per launch = 2 * B*T * 768 * 256 (padded frames included: the reference computes them too)."""
    from pytorch_kaldi_asr_b200 import ops
    layer = model.encoder_test.tdnn_stack[2]
    ctx = layer.concat.index
    flops = 2.0 * B * T * 768 * 256
    pk_ = peaks()
    out = {}
    if mode == "bf16":
        xs = [torch.randn(B, T, 256, device="cuda").bfloat16() for _ in range(24)]          # 24 x 8 MB > 126 MB L2
        dzs = [torch.randn(B, T, 256, device="cuda").bfloat16() for _ in range(24)]
        wf, wd = ops.weight_relayout(layer.proj.weight.detach(), 256, 3)
        bias = layer.proj.bias.detach()
        i = [0]

        def fwd():
            i[0] += 1
            ops.gemm_tc_rows(xs[i[0] % 24], wf, B, T, 256, 256, nseg=3, lda=256, ldb=768, b_seg_col=256, shift=ctx, bias=bias, relu=True)

        def dgrad():
            i[0] += 1
            ops.gemm_tc_rows(dzs[i[0] % 24], wd, B, T, 256, 256, nseg=3, lda=256, ldb=768, b_seg_col=256, shift=[-c for c in ctx])

        def wgrad():
            i[0] += 1
            ops.gemm_tc_wgrad(dzs[i[0] % 24], xs[i[0] % 24], B, T, 256, 256, 3, ctx)

        secs = {"fwd": time_kernel(fwd), "dgrad": time_kernel(dgrad), "wgrad": time_kernel(wgrad)}
        sec = secs["fwd"]
        out["kernel"] = ("gemm_tc_rows2_kernel<pair> (tcgen05.mma cta_group::2 256x256x16, TMEM accumulators, activations resident "
                         "with splice halo, split weight tiles by TMA, TMA-store epilogue; spliced TDNN forward)")
        out["family_us"] = {k: v * 1e6 for k, v in secs.items()}
        out["family_tflops"] = {k: flops / v / 1e12 for k, v in secs.items()}
    else:
        xs = [torch.randn(B, T, 256, device="cuda") for _ in range(12)]
        i = [0]

        def fwd():
            with torch.no_grad():
                ops.linear(xs[i[0] % 12], layer.proj.weight, layer.proj.bias, splice=ctx, relu=True)
            i[0] += 1

        sec = time_kernel(fwd)
        out["kernel"] = "gemm_f32_kernel<128,128,32,8,8> (TDNN splice+Linear+bias+ReLU, fp32 SIMT exact path)"
    achieved = flops / sec / 1e12
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")      # DRAM bytes per launch from the committed ncu capture
    if mode == "bf16" and os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    out.update({"bound": "tensor", "achieved": achieved, "peak": pk_["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk_["bf16_tflops"], "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": pk_["source"] + " bf16 burst",
                "us_per_launch": sec * 1e6, "flops_per_launch": flops})
    return out


def hbm_probe():
    from pytorch_kaldi_asr_b200 import ops
    rows, D = 32 * 430 * 16, 256                      # 225 MB per tensor > L2
    x = torch.randn(rows, D, device="cuda")
    r = torch.randn(rows, D, device="cuda")
    a, b = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")

    def f():
        with torch.no_grad():
            ops.add_layer_norm(x.view(1, rows, D), r.view(1, rows, D), a, b)

    sec = time_kernel(f, iters=4, replays=3)
    pk_ = peaks()
    gbs = 3.0 * rows * D * 4 / sec / 1e9
    return {"kernel": "add_ln_fwd_kernel", "achieved": gbs, "peak": pk_["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk_["hbm_gbs"],
            "bytes_per_launch": 3 * rows * D * 4,
            "note": "2 reads : 1 write per element; the peak is the measured 1:1 copy figure, which a read-heavy kernel "
                    "can exceed by a few percent (ncu: 98 % of it in DRAM throughput, profiles/r01_kernel_evidence.md)"}


def cfg5_bench(B=4, band=(-100, 0), steps=5, warmup=3, profile=False):
    """BASELINE config 5 (the attention stress case): self-attention `Encoder` 12 layers + `Decoder` 6 layers, d_model 512,
    H=8, d_k=d_v=64, ~1500-frame utterances (T_i = clip(N(1500,100),1200,1599), L_i = T_i//10), bf16 tensor-core path,
    one training step = forward + summed CE + backward + Adam, replayed as one CUDA graph."""
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200.utils import synthetic
    pk.set_compute_mode("bf16")
    cfg = dict(n_src_dim=40, n_tgt_vocab=53, encoder_max_len=1600, decoder_max_len=200, src_fold=1,
               encoder_sub_sequence=band, decoder_sub_sequence=(-20, 0), en_layers=12, de_layers=6, n_head=8,
               en_d_model=512, de_d_model=512, d_k=64, d_v=64, en_dropout=0.1, de_dropout=0.1)
    torch.manual_seed(0)
    model = pk.Transformer(lda_mat=None, encoder_type="attention", **cfg).cuda()
    opt = pk.ScheduledOptim(pk.FusedAdam(model.parameters(), betas=(0.9, 0.999), eps=1e-8), 1e-3, 25000)
    pool = synthetic.batches(4, B, seed=555, pad_to="set", mean_len=1500.0, std_len=100.0, min_len=1200, max_len=1599,
                             label_div=10, max_labels=198)
    dev_pool = [pk.train._to_device(b, "cuda", non_blocking=False) for b in pool]
    frames = [synthetic.real_frames(b) for b in pool]
    model.train()
    graphed = pk.GraphedTrainStep(model, opt, pool[0])        # all batches are padded to one shape -> one CUDA graph

    def step(i):
        graphed.load(*dev_pool[i % len(dev_pool)])
        graphed.graph.replay()
        return graphed.out[0]

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        loss = step(warmup + i)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    fr = float(np.mean([frames[(warmup + i) % len(frames)] for i in range(steps)]))
    n_par = sum(p.numel() for p in model.parameters() if p.requires_grad)
    out = {"workload": "cfg5: Encoder 12L + Decoder 6L, d_model 512, H=8, T~1500 (BASELINE configs[4])", "batch": B,
           "band": list(band), "padded_T": int(pool[0][1].shape[1]), "padded_L": int(pool[0][3].shape[1] - 1),
           "ms_per_step": ms, "frames_per_sec": fr / (ms * 1e-3), "utts_per_sec": B / (ms * 1e-3), "trainable_params": int(n_par),
           "loss_finite": bool(torch.isfinite(loss).item()), "dtype": "bf16", "execution": "cuda-graph"}
    if profile:
        import collections, re
        from torch.profiler import profile as tprofile, ProfilerActivity
        with tprofile(activities=[ProfilerActivity.CUDA]) as prof:
            step(0)
            torch.cuda.synchronize()
        agg = collections.defaultdict(lambda: [0, 0.0])
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                name = re.sub(r"\(.*", "", ev.name)[:70]
                agg[name][0] += 1
                agg[name][1] += ev.device_time_total
        tot = sum(v[1] for v in agg.values())
        out["kernel_us_per_step"] = tot
        out["top_kernels"] = [(k, round(v[1], 1), v[0]) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]]
    pk.set_compute_mode("bf16")
    return out


def decode_bench(model, pk, rank, world, n_total=1000, batch=125, beam=10, max_len=100):
    """BASELINE config 4: beam-10 decoding of 1000 synthetic utterances, sharded over the ranks with no collective.
    fp32 exact path.  Returns per-rank (seconds_forced, seconds_natural, n_utts, steps_natural)."""
    import types
    from pytorch_kaldi_asr_b200 import parallel
    from pytorch_kaldi_asr_b200.decode import translate_batch
    from pytorch_kaldi_asr_b200.utils import synthetic
    from pytorch_kaldi_asr_b200.utils.instances_handler import pad_to_longest
    pk.set_compute_mode("fp32")
    feats, _ = synthetic.utterances(n_total, np.random.RandomState(4321))
    lo, hi = parallel.shard_range(n_total, rank, world)
    src_all, mask_all = pad_to_longest(feats[lo:hi])              # one padded length per shard -> one step graph
    batches = []
    for i in range(0, hi - lo, batch):
        if i + batch > hi - lo:
            i = max(0, hi - lo - batch)                           # last batch re-decodes a few utterances, same shape
        batches.append((None, torch.from_numpy(src_all[i:i + batch]).pin_memory(),
                        torch.from_numpy(mask_all[i:i + batch]).pin_memory(), None, None))
    out = {}
    for forced in (True, False):
        opt = types.SimpleNamespace(use_gpu=True, beam_size=beam, max_token_seq_len=max_len, nbest=1, force_full_length=forced)
        translate_batch(model, batches[0], opt, None)             # warm-up: buffers + step graph for the first shape
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for b in batches:
            translate_batch(model, b, opt, None)                  # ends with the D2H read-out of the lattices
        torch.cuda.synchronize()
        out["forced" if forced else "natural"] = time.perf_counter() - t0
    return out, hi - lo


def cpu_decode_baseline(n_utt, beam, max_len, threads):
    from oracle import acoustic_model as am
    from oracle import beam_decode as obd
    from pytorch_kaldi_asr_b200.utils import synthetic
    from pytorch_kaldi_asr_b200.utils.instances_handler import pad_to_longest
    torch.set_num_threads(threads)
    cfg = am.example_config()
    sd = am.init_state_dict(cfg, synthetic.lda_matrix(), seed=0)
    feats, _ = synthetic.utterances(n_utt, np.random.RandomState(4321))
    src, mask = pad_to_longest(feats)
    t0 = time.perf_counter()
    obd.translate_batch(sd, cfg, src, mask, beam, max_len, 1, force_full_length=True)
    return n_utt / (time.perf_counter() - t0)


def run_b200(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (B200 arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # communicator creation prints an "NCCL version" banner on stdout: route fd 1 to stderr until the first
        # collective is through, so that stdout carries exactly ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            if rank == 0:
                ensure_library()
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    elif rank == 0:
        ensure_library()
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200 import _lib, parallel
    from pytorch_kaldi_asr_b200.utils import synthetic
    _lib.check(_lib.lib().pka_check_device(), "check_device")
    pk.set_compute_mode(args.mode)

    B = args.batch
    cfg = dict(n_src_dim=40, n_tgt_vocab=53, encoder_max_len=500, decoder_max_len=100, src_fold=1,
               encoder_sub_sequence=(-100, 0), decoder_sub_sequence=(-10, 0), en_layers=3, de_layers=3, n_head=2,
               en_d_model=256, de_d_model=128, d_k=64, d_v=64, en_dropout=0.35, de_dropout=0.35,
               tdnn_contexts=[[-1, 0, 1], [-1, 0, 1], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3]])
    torch.manual_seed(0)                                         # identical replicas on every rank
    model = pk.Transformer(lda_mat=synthetic.lda_matrix(), seed=1000 + rank, **cfg).cuda()
    opt = pk.ScheduledOptim(pk.FusedAdam(model.parameters(), betas=(0.9, 0.999), eps=1e-8), 1e-3, 25000)
    n_pool = 8
    pool = synthetic.batches(n_pool, B, seed=1234 + 7919 * rank, pad_to="set")     # every rank its own utterances
    frames_pool = [synthetic.real_frames(b) for b in pool]
    dev_pool = [pk.train._to_device(b, "cuda", non_blocking=False) for b in pool]
    pin_pool = [(None,) + tuple(torch.as_tensor(np.ascontiguousarray(x)).to(dt).pin_memory()
                                for x, dt in zip(b[1:], (torch.float32, torch.uint8, torch.int64, torch.uint8)))
                for b in pool]
    n_buckets = int(os.environ.get("PKA_BUCKETS", "3"))
    dp_backend = os.environ.get("PKA_ALLREDUCE", "nccl")
    if world > 1 and dp_backend == "peer":
        # one kernel per rank: gradient reduce-scatter + Adam on the shard + parameter all-gather over peer memory
        opt.optimizer.enable_peer_step()
        sync = None
    else:
        sync = parallel.GradAllReduce(opt.optimizer, n_buckets=n_buckets, backend=dp_backend) if world > 1 else None
    launches0 = _lib.launch_count()
    graphed, graph_note = None, "eager"
    if not args.no_graph:
        try:
            graphed = pk.GraphedTrainStep(model, opt, pool[0], grad_sync=sync.finish if sync else None)
            graph_note = "cuda-graph"
        except Exception as exc:                                  # stay on the GPU path, just without the graph
            if world == 1:
                raise
            graph_note = "eager (graph capture with NCCL failed: %s)" % str(exc).split("\n")[0][:80]
            graphed = None
            torch.cuda.synchronize()
    launches_per_step = (_lib.launch_count() - launches0) // 4 if graphed is not None else None   # 3 warm-up + 1 capture

    class Loader(list):
        mode = "drop"

    def eager_step(batch_dev):
        src, smask, tgt, tmask = batch_dev
        opt.zero_grad()
        pred = model(src, smask, tgt[:, :-1], tmask[:, :-1])
        loss, stats = pk.ops.cross_entropy_sum(pred.view(-1, pred.size(-1)), tgt[:, 1:].contiguous().view(-1), False)
        loss.backward()
        if sync is not None:
            sync.finish()
        opt.step()
        opt.update_learning_rate()
        return loss

    def resident_step(i):
        if graphed is not None:
            graphed.load(*dev_pool[i % n_pool])
            graphed.graph.replay()
        else:
            eager_step(dev_pool[i % n_pool])

    def e2e_run(first, count):
        """The public API call a user makes: one train_epoch over `count` pinned host batches; every step copies its
        inputs host->device (overlapped with the previous step's kernels) and reads loss / accuracy back to the host."""
        loader = Loader([pin_pool[(first + i) % n_pool] for i in range(count)])
        if graphed is not None or sync is None:
            return pk.train_epoch(model, loader, None, mode="train", optimizer=opt, graphed=graphed, sync_every_step=True)
        for b in loader:
            float(eager_step(pk.train._to_device(b, "cuda")))     # D2H read of the loss every step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, run_fn=None):
        if run_fn is not None:
            run_fn(0, args.warmup)
        else:
            for i in range(args.warmup):
                step_fn(i)
        barrier()
        import gc
        gc.collect()
        gc.disable()                              # no collector pauses inside the timed region
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        s.record()
        if run_fn is not None:
            run_fn(args.warmup, args.steps)
        else:
            for i in range(args.steps):
                step_fn(args.warmup + i)
        e.record()
        barrier()
        gc.enable()
        ms = s.elapsed_time(e)
        frames = sum(frames_pool[(args.warmup + i) % n_pool] for i in range(args.steps))
        t = torch.tensor([ms, float(frames)], device="cuda", dtype=torch.float64)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms, frames = float(tmax[0]), float(t[1])
        return ms, frames, _lib.launch_count() - n0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    ms_res, frames_res, eager_launches = timed(resident_step)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, frames_e2e, _ = timed(None, e2e_run)

    dec = None
    if not args.no_decode:
        dec_t, dec_n = decode_bench(model, pk, rank, world)
        t = torch.tensor([dec_t["forced"], dec_t["natural"]], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dec = {"metric": "decode_utts_per_sec", "unit": "utts/s", "n_utts": 1000, "beam": 10, "max_token_seq_len": 100,
               "value_forced_100_steps": 1000.0 / float(t[0]), "value_natural_eos": 1000.0 / float(t[1]),
               "dtype": "f32", "sharding": "%d utterances per rank, no collective" % dec_n,
               "timing": "wall clock around translate_batch over the rank's shard incl. H2D of features and D2H of lattices, max over ranks"}
    if rank != 0:
        _finish(world)
        return
    value = frames_res / (ms_res * 1e-3)
    e2e_value = frames_e2e / (ms_e2e * 1e-3)
    b0 = pool[0]
    h2d = sum(int(np.asarray(x).nbytes) for x in b0[1:] if x is not None)
    h2d = int(b0[1].astype(np.float32).nbytes + b0[2].astype(np.uint8).nbytes + b0[3].astype(np.int64).nbytes + b0[4].astype(np.uint8).nbytes)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.mode == "fp32" else "bf16", "data": "synthetic",
        "config": {"workload": "timit_example_model_train_step_B32 (BASELINE configs[1])", "global_batch": B * world,
                   "batch_per_gpu": B, "padded_T": int(b0[1].shape[1]), "padded_L": int(b0[3].shape[1] - 1),
                   "dropout": 0.35, "parallelism": "dp%d" % world, "execution": graph_note,
                   "gradient_exchange": "none (1 GPU)" if world == 1 else dp_backend,
                   "l2": "no flush: %d distinct batches rotate and one step's activation working set (~0.5 GB) exceeds the 126 MB L2" % n_pool},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
                "ms_per_step": ms_e2e / args.steps, "api": "train_epoch(model, loader_of_K_pinned_batches, None, 'train', optimizer, graphed=..., sync_every_step=True)"},
        "gpu_launches": int(launches_per_step * args.steps) if launches_per_step else int(eager_launches),
        "launches_per_step": launches_per_step,
        "clocks": clocks,
        "real_frames_per_step_per_gpu": float(np.mean(frames_pool)),
    }
    if dec is not None:
        line["decode"] = dec
    if world == 1:
        if dec is not None:
            try:
                v = cpu_decode_baseline(4, 10, 100, os.cpu_count() or 1)
                dec["cpu_baseline"] = {"value": v, "unit": "utts/s", "cores": os.cpu_count() or 1, "kind": "port",
                                       "sample": "4 utterances, beam 10, 100 forced steps, oracle port of L/decode.py:22-107 (no KV cache)"}
            except Exception as exc:
                dec["cpu_baseline"] = {"error": str(exc)[:200]}
        try:
            line["roofline"] = roofline_probe(model, B, int(b0[1].shape[1]), pk, args.mode)
            line["roofline_hbm"] = hbm_probe()
        except Exception as exc:                                   # never lose the headline line to a probe
            line["roofline"] = {"error": str(exc)[:200]}
        if not args.no_cfg5:
            try:
                line["cfg5"] = cfg5_bench(B=8, band=(-100, 0), steps=5, warmup=3)
                line["cfg5_full_attention"] = cfg5_bench(B=4, band=(-1600, 1600), steps=3, warmup=2)
                pk.set_compute_mode(args.mode)
            except Exception as exc:
                line["cfg5"] = {"error": str(exc)[:200]}
        threads = os.cpu_count() or 1
        t0 = time.time()
        v, ms = cpu_train_baseline(32, 8, 1, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_step": ms,
                                "sample": "8 steps x 32 utterances (same synthetic batches), oracle port of the reference "
                                          "train step, dropout 0.35, %.0f s of CPU work" % (time.time() - t0)}
    print(json.dumps(line))
    _finish(world)


def _finish(world):
    """Leave without tearing NCCL down: destroying a process group whose collectives were captured into CUDA graphs can
    block forever at exit, so after a last barrier every rank flushes and exits hard."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--mode", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
