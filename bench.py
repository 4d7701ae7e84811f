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
import math
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


# ---------------------------------------------------------------------------------------------------- the workload
WORKLOAD = "timit_example_model_train_step_B32 (BASELINE configs[1])"
N_POOL = 8


def model_config():
    return dict(n_src_dim=40, n_tgt_vocab=53, encoder_max_len=500, decoder_max_len=100, src_fold=1,
                encoder_sub_sequence=(-100, 0), decoder_sub_sequence=(-10, 0), en_layers=3, de_layers=3, n_head=2,
                en_d_model=256, de_d_model=128, d_k=64, d_v=64, en_dropout=0.35, de_dropout=0.35,
                tdnn_contexts=[[-1, 0, 1], [-1, 0, 1], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3]])


def build_pool(B, rank, padding):
    """The synthetic batches BOTH arms step through: N_POOL batches of B utterances, seed 1234 on rank 0
    (+7919 per rank).  padding "set": every batch padded to the longest utterance of the set, the reference loader's
    behaviour (U/BatchLoader.py:33-36).  padding "bucket": utterances sorted by length into batches, each batch padded
    to min(ceil((longest + 16) / 64) * 64, set length) -- the TDNN stack's right context is 16 frames, so every real
    frame sees exactly what it sees under whole-set padding (tests/test_gpu_bucketed.py) -- labels to a multiple of 8."""
    from pytorch_kaldi_asr_b200.utils import synthetic
    if padding == "set":
        return synthetic.batches(N_POOL, B, seed=1234 + 7919 * rank, pad_to="set")
    return synthetic.batches(N_POOL, B, seed=1234 + 7919 * rank, pad_to="bucket")


def workload_config(B, world, pool, padding):
    """`config` of the JSON line -- identical in both arms by construction."""
    from pytorch_kaldi_asr_b200.utils import synthetic
    Ts = sorted({int(b[1].shape[1]) for b in pool})
    Ls = sorted({int(b[3].shape[1] - 1) for b in pool})
    real = float(np.mean([synthetic.real_frames(b) for b in pool]))
    padded = float(np.mean([b[1].shape[0] * b[1].shape[1] for b in pool]))
    return {"workload": WORKLOAD, "global_batch": B * world, "batch_per_gpu": B, "padding": padding,
            "padded_T": Ts if len(Ts) > 1 else Ts[0], "padded_L": Ls if len(Ls) > 1 else Ls[0],
            "real_frames_per_batch": real, "padded_frames_per_batch": padded, "real_fraction": real / padded,
            "dropout": 0.35, "n_distinct_batches": len(pool), "parallelism": "dp%d" % world,
            "l2": "no flush: %d distinct batches rotate and one step's activation working set (~0.5 GB) exceeds the "
                  "126 MB L2" % len(pool)}


# ---------------------------------------------------------------------------------------------------- CPU arm (oracle)
def cpu_train_baseline(pool, steps: int, warmup: int, threads: int):
    """The reference algorithm (oracle port: torch-CPU fp32, dropout 0.35 through torch's generator like nn.Dropout)
    timed on the host cores: forward + loss + backward + Adam + LR tick, one batch of `pool` per step (the same batches,
    in the same order, as the B200 arm)."""
    from oracle import acoustic_model as am
    from oracle import train_step as otrain
    from pytorch_kaldi_asr_b200.utils import synthetic
    torch.set_num_threads(threads)
    cfg = am.example_config()
    sd = am.init_state_dict(cfg, synthetic.lda_matrix(), seed=0)
    opt = otrain.AdamSchedule({k: sd[k] for k in am.trainable_keys(sd)}, 1e-3, 25000)

    def one(batch):
        _, loss, _, _, grads = otrain.loss_and_grads(sd, cfg, batch[1:], False, am.DropoutPlan("rng"))
        opt.step(grads)
        opt.update_learning_rate()
        return float(loss)

    for i in range(warmup):
        one(pool[i % len(pool)])
    frames, t0 = 0, time.perf_counter()
    for i in range(steps):
        one(pool[(warmup + i) % len(pool)])
        frames += synthetic.real_frames(pool[(warmup + i) % len(pool)])
    dt = time.perf_counter() - t0
    return frames / dt, dt / steps * 1e3


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port; /root/reference does not travel
    to the GPU box) on all host threads, same batches / padding / config / steps / warm-up as the B200 arm.  Each step
    is one rank's share (B utterances) of the workload -- the bounded sample of a dp-N global batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = args.batch
    pool = build_pool(B, 0, args.padding)
    value, ms = cpu_train_baseline(pool, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(B, args.gpus, pool, args.padding),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d steps x %d utterances (one rank's batches of the workload), oracle port of "
                                   "L/train.py:145-207, fp32, dropout 0.35" % (args.steps, B)},
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
    """The residual-add + LayerNorm forward kernel THE TIMED STEP RUNS (bf16 activation stream, D = 128:
    add_ln_fwd_kernel<__nv_bfloat16, 8, 16, 1, DROP>: 16-byte vectors, 16 lanes per row = two rows per warp iteration,
    persistent grid with the gain / offset in registers), at the step's own shape (B*L = 2016 rows of 128: 1.5 MB, latency-bound) and
    at a bandwidth-sized shape (> L2), CUDA-graph-timed alone.  Algorithmic bytes = 3 * rows * D * 2 (read x, read
    residual, write y)."""
    from pytorch_kaldi_asr_b200 import ops
    pk_ = peaks()
    out = {"kernel": "add_ln_fwd_kernel<__nv_bfloat16, 8, 16, 1, false> (dropout(x) + residual -> LayerNormalization, bf16 in/out; "
                     "16 lanes x 16 bytes per 128-wide row, persistent grid, next rows prefetched)",
           "bound": "hbm", "peak": pk_["hbm_gbs"], "unit": "GB/s", "peak_source": pk_["source"] + " copy bandwidth"}
    for name, rows, D, nbuf in (("in_step", 2016, 128, 64), ("bandwidth_sized", 32 * 430 * 64, 128, 2)):
        xs = [torch.randn(rows, D, device="cuda").bfloat16() for _ in range(nbuf)]      # in-step: 64 x 0.5 MB rotate
        rs = [torch.randn(rows, D, device="cuda").bfloat16() for _ in range(nbuf)]
        a, b = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
        i = [0]

        def f():
            i[0] += 1
            with torch.no_grad():
                ops.add_layer_norm(xs[i[0] % nbuf].view(1, rows, D), rs[i[0] % nbuf].view(1, rows, D), a, b)

        sec = time_kernel(f, iters=16 if rows < 100000 else 4, replays=5 if rows < 100000 else 3)
        nbytes = 3.0 * rows * D * 2
        out[name] = {"rows": rows, "D": D, "bytes_per_launch": nbytes, "us_per_launch": sec * 1e6,
                     "achieved": nbytes / sec / 1e9, "frac": nbytes / sec / 1e9 / pk_["hbm_gbs"]}
        del xs, rs
    out["achieved"], out["frac"] = out["bandwidth_sized"]["achieved"], out["bandwidth_sized"]["frac"]
    out["note"] = ("frac/achieved = the bandwidth-sized shape; in_step is the 2016 x 128 launch of the timed step, whose 1.5 MB "
                   "cannot reach HBM bandwidth in one ~3 us launch (latency-bound)")
    return out


# ---------------------------------------------------------------------------------------------------- algorithmic FLOPs
def _band_pairs(n_q, n_k, band):
    """Allowed (query, key) pairs: key j real (< n_k) and i+start <= j <= i+end (T/Models.py:38-49)."""
    if band is None:
        return n_q * n_k
    i = np.arange(n_q)
    lo = np.maximum(0, i + band[0])
    hi = np.minimum(n_k - 1, i + band[1])
    return int(np.maximum(0, hi - lo + 1).sum())


def decoder_fwd_flops(T, L, Dd, De, H, dk, Nd, V, band):
    """SURVEY Appendix C, forward FLOPs (2 per MAC) of the decoder for one utterance with T real frames, L real tokens."""
    per_pair = 4 * H * dk
    per_layer = L * (8 * Dd * H * dk + 4 * H * dk * Dd + 4 * Dd * Dd) + _band_pairs(L, L, band) * per_pair \
        + L * T * per_pair + T * 4 * Dd * H * dk
    return Nd * per_layer + T * 2 * De * Dd + L * 2 * Dd * V


def timit_step_flops(batch):
    """Algorithmic fwd+bwd FLOPs of one TIMIT-config train step on REAL frames / tokens only."""
    lens = np.asarray(batch[2]).sum(axis=1)
    toks = np.asarray(batch[4])[:, :-1].sum(axis=1)
    enc = 7362688.0 * lens.sum()
    dec = sum(3.0 * decoder_fwd_flops(int(t), int(l), 128, 256, 2, 64, 3, 53, (-10, 0)) for t, l in zip(lens, toks))
    return enc + dec


def cfg5_step_flops(batch, band, D=512, H=8, dk=64, Ne=12, Nd=6, F=40, V=53, dband=(-20, 0)):
    lens = np.asarray(batch[2]).sum(axis=1)
    toks = np.asarray(batch[4])[:, :-1].sum(axis=1)
    total = 0.0
    for t, l in zip(lens, toks):
        t, l = int(t), int(l)
        enc = 2 * t * F * D + Ne * (t * (8 * D * H * dk + 4 * D * D) + _band_pairs(t, t, band) * 4 * H * dk)
        total += 3.0 * (enc + decoder_fwd_flops(t, l, D, D, H, dk, Nd, V, dband))
    return total


def cfg5_bench(B=4, band=(-100, 0), steps=5, warmup=3, profile=False):
    """BASELINE config 5 (the attention stress case): self-attention `Encoder` 12 layers + `Decoder` 6 layers, d_model 512,
    H=8, d_k=d_v=64, ~1500-frame utterances (T_i = clip(N(1500,100),1200,1599), L_i = T_i//10), bf16 tensor-core path,
    one training step = forward + summed CE + backward + Adam, replayed as one CUDA graph.  `roofline`: algorithmic
    FLOPs of the step (real positions, allowed pairs; SURVEY Appendix C, backward = 2 x forward) / step time against the
    SUSTAINED bf16 peak (a kernel timed inside a long step)."""
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
    flops = [cfg5_step_flops(b, band) for b in pool]
    model.train()
    graphed = pk.GraphedTrainStep(model, opt, pool[0])        # all batches are padded to one shape -> one CUDA graph

    def step(i):
        return graphed.step(*dev_pool[i % len(dev_pool)])

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        out3 = step(warmup + i)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    fr = float(np.mean([frames[(warmup + i) % len(frames)] for i in range(steps)]))
    fl = float(np.mean([flops[(warmup + i) % len(flops)] for i in range(steps)]))
    n_par = sum(p.numel() for p in model.parameters() if p.requires_grad)
    pk_ = peaks()
    tf = fl / (ms * 1e-3) / 1e12
    out = {"workload": "cfg5: Encoder 12L + Decoder 6L, d_model 512, H=8, T~1500 (BASELINE configs[4])", "batch": B,
           "band": list(band), "padded_T": int(pool[0][1].shape[1]), "padded_L": int(pool[0][3].shape[1] - 1),
           "ms_per_step": ms, "frames_per_sec": fr / (ms * 1e-3), "utts_per_sec": B / (ms * 1e-3), "trainable_params": int(n_par),
           "loss_finite": bool(torch.isfinite(out3[0]).item()), "dtype": "bf16", "execution": "cuda-graph",
           "parity": "tests/test_gpu_bf16_at_size.py (this model at this size against the fp32 oracle)",
           "roofline": {"bound": "tensor", "achieved": tf, "peak": pk_["bf16_tflops_sustained"], "unit": "TFLOP/s",
                        "frac": tf / pk_["bf16_tflops_sustained"], "flops_per_step": fl,
                        "peak_source": pk_["source"] + " bf16 sustained",
                        "convention": "whole train step; algorithmic FLOPs on real positions and allowed pairs, bwd = 2 x fwd"}}
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
    graphed.release()
    return out


def attention_probe(T=1600, B=4, H=8, band=None):
    """The kernel config 5 leans on: attn_tc_fwd_kernel at T = 1600, d = 64, timed alone.  Algorithmic FLOPs = allowed
    pairs x 4 x 64 per head (QK^T and PV, 2 per MAC), against the bf16 BURST peak (kernel timed alone)."""
    from pytorch_kaldi_asr_b200 import ops
    HD = H * 64
    qkv = [(torch.randn(B, T, 3 * HD, device="cuda") * 0.5).bfloat16() for _ in range(4)]
    mask = torch.ones(B, T, device="cuda", dtype=torch.uint8)
    i = [0]

    def f():
        i[0] += 1
        with torch.no_grad():
            ops.attention_tc(qkv[i[0] % 4], None, mask, H, 64, band, 1.0 / math.sqrt(512.0), None)

    sec = time_kernel(f, iters=8, replays=3)
    flops = B * H * _band_pairs(T, T, band) * 4.0 * 64
    pk_ = peaks()
    return {"kernel": "attn_tc_fwd_kernel (tcgen05 QK^T / PV, TMEM, TMA, online softmax)", "B": B, "T": T, "H": H,
            "band": list(band) if band else None, "us_per_launch": sec * 1e6, "flops_per_launch": flops,
            "achieved": flops / sec / 1e12, "peak": pk_["bf16_tflops"], "unit": "TFLOP/s", "bound": "tensor",
            "frac": flops / sec / 1e12 / pk_["bf16_tflops"]}


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


def decode_roofline_probe(n_utt=125, T=499, beam=10, H=2, dk=64):
    """Dominant HBM stream of a beam-search step: cross-attention of the `beam` live hypotheses of every utterance over
    its T encoder frames (attn_fwd_smallq_kernel, fp32 exact path).  Algorithmic bytes per launch = the K/V rows read
    once per utterance, n_utt * T * (dk + dv) * H * 4; 3 launches per step (one per decoder layer)."""
    from pytorch_kaldi_asr_b200 import ops
    HD = H * dk
    nbuf = 4                                                      # 4 x 64 MB of K/V rotate (> L2)
    kvs = [torch.randn(n_utt, T, 2 * HD, device="cuda") for _ in range(nbuf)]
    q = torch.randn(n_utt, beam, HD, device="cuda")
    mask = torch.ones(n_utt, T, device="cuda", dtype=torch.uint8)
    i = [0]

    def f():
        i[0] += 1
        with torch.no_grad():
            ops.attention(q, kvs[i[0] % nbuf], mask, H, dk, None, 1.0 / math.sqrt(128.0), None)

    sec = time_kernel(f, iters=8, replays=3)
    nbytes = float(n_utt) * T * 2 * HD * 4
    pk_ = peaks()
    return {"kernel": "attn_fwd_smallq_kernel (beam = query axis, K/V shared per utterance, fp32)", "bound": "hbm",
            "n_utt": n_utt, "T": T, "beam": beam, "bytes_per_launch": nbytes, "us_per_launch": sec * 1e6,
            "achieved": nbytes / sec / 1e9, "peak": pk_["hbm_gbs"], "unit": "GB/s", "frac": nbytes / sec / 1e9 / pk_["hbm_gbs"],
            "peak_source": pk_["source"] + " copy bandwidth"}


def cpu_decode_baseline(n_utt, beam, max_len, threads):
    from oracle import acoustic_model as am
    from oracle import beam_decode as obd
    from pytorch_kaldi_asr_b200.utils import synthetic
    from pytorch_kaldi_asr_b200.utils.instances_handler import pad_to_longest
    torch.set_num_threads(threads)
    cfg = am.example_config()
    sd = am.init_state_dict(cfg, synthetic.lda_matrix(), seed=0)
    feats, _ = synthetic.utterances(n_utt, np.random.RandomState(4321))
    t0 = time.perf_counter()
    for i in range(0, n_utt, 8):                                  # batches of 8 like the reference recipe (P/run.sh:160)
        src, mask = pad_to_longest(feats[i:i + 8])
        obd.translate_batch(sd, cfg, src, mask, beam, max_len, 1, force_full_length=True)
    return n_utt / (time.perf_counter() - t0)


# ---------------------------------------------------------------------------------------------------- DP parity (N > 1)
def dp_parity_check(model, opt, sync, pool, rank, world, pk, mode):
    """Assertions the scaling runs carry (SURVEY section 4 'distributed' level), run after the timed blocks:
      (1) every rank holds bit-identical parameters after the data-parallel steps;
      (2) one DP backward's all-reduced gradient == the single-process gradient of the concatenated N*B batch padded to
          one global T (SUM semantics, L/train.py:86-88), dropout off so both sides see the same function;
      (3) sharded decoding (each rank its contiguous shard, no collective) == unsharded decoding of the same utterances
          (L/decode.py:154-161), token for token."""
    import types
    import torch.distributed as dist
    from pytorch_kaldi_asr_b200 import parallel
    from pytorch_kaldi_asr_b200.decode import translate_batch
    from pytorch_kaldi_asr_b200.utils import synthetic
    from pytorch_kaldi_asr_b200.utils.instances_handler import pad_to_longest
    inner = opt.optimizer
    out = {}
    # (1)
    hi, lo = inner.flat_param.clone(), inner.flat_param.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    out["params_bit_identical"] = bool(torch.equal(hi, lo))
    # (2) every rank's first batch, re-padded to one global shape
    b = pool[0]
    shape = torch.tensor([b[1].shape[1], b[3].shape[1]], device="cuda")
    dist.all_reduce(shape, op=dist.ReduceOp.MAX)
    Tg, Lg = int(shape[0]), int(shape[1])

    def repad(x, n, fill=0):
        out_ = np.full((x.shape[0], n) + x.shape[2:], fill, dtype=x.dtype)
        out_[:, :x.shape[1]] = x
        return out_

    mine = (None, repad(b[1], Tg), repad(b[2], Tg), repad(b[3], Lg), repad(b[4], Lg))
    dev = pk.train._to_device(mine, "cuda", non_blocking=False)
    model.eval()                                                   # dropout off; nothing else depends on the mode

    def backward(batch_dev):
        src, smask, tgt, tmask = batch_dev
        opt.zero_grad()
        pred = model(src, smask, tgt[:, :-1], tmask[:, :-1])
        loss, _ = pk.ops.cross_entropy_sum(pred.view(-1, pred.size(-1)), tgt[:, 1:].contiguous().view(-1), False)
        loss.backward()
        return loss

    loss = backward(dev)
    if sync is not None:
        sync.finish()
    inner._adopt_grads()
    g_dp = inner.flat_grad.clone()
    if sync is None:
        # peer mode: the gradients are summed inside the fused optimiser kernel, never in place.  Sum a copy with NCCL
        # for the comparison below, and hold the kernel itself to {NCCL SUM, local Adam} on the same state.
        import ctypes as C
        from pytorch_kaldi_asr_b200 import _lib as L
        dist.all_reduce(g_dp, op=dist.ReduceOp.SUM)
        inner.sync_moments()
        state = [inner.flat_param, inner.exp_avg, inner.exp_avg_sq, inner.dev_state, inner.dev_lr, inner.flat_grad]
        snap = [t.clone() for t in state]
        inner.step()                                               # the peer kernel (collective)
        torch.cuda.synchronize()
        dist.barrier()
        after_peer = inner.flat_param.clone()
        for t, s_ in zip(state, snap):
            t.copy_(s_)
        torch.cuda.synchronize()
        dist.barrier()
        g = inner.param_groups[0]
        L.check(L.lib().pka_adam_step(L.ptr(inner.flat_param), L.ptr(g_dp), L.ptr(inner.exp_avg), L.ptr(inner.exp_avg_sq),
                                      C.c_int64(inner.numel), L.ptr(inner.dev_lr), C.c_float(g["lr"]), L.ptr(inner.dev_state),
                                      C.c_float(g["betas"][0]), C.c_float(g["betas"][1]), C.c_float(g["eps"]), C.c_void_p(0),
                                      L.stream_ptr()), "adam_step")
        torch.cuda.synchronize()
        out["peer_step_max_abs_diff_vs_nccl_sum_plus_adam"] = float((after_peer - inner.flat_param).abs().max())
        out["peer_step_bit_identical"] = bool(torch.equal(after_peer, inner.flat_param))
        for t, s_ in zip(state, snap):
            t.copy_(s_)
        torch.cuda.synchronize()
        dist.barrier()
    loss_sum = loss.detach().clone().double()
    dist.all_reduce(loss_sum)
    gathered = [[torch.empty_like(t) for _ in range(world)] for t in dev]
    for t, lst in zip(dev, gathered):
        dist.all_gather(lst, t.contiguous())
    if rank == 0:
        cat = tuple(torch.cat(lst, dim=0) for lst in gathered)
        ctx = sync.paused() if sync is not None else None
        if ctx is not None:
            with ctx:
                loss1 = backward(cat)
        else:
            loss1 = backward(cat)
        inner._adopt_grads()
        g_one = inner.flat_grad
        worst = 0.0
        for p, off in zip(inner._train, inner._offsets):
            a, r = g_dp[off:off + p.numel()], g_one[off:off + p.numel()]
            worst = max(worst, float((a - r).abs().max() / r.abs().max().clamp_min(1e-12)))
        out["grad_rel_err_vs_single_process"] = worst
        out["grad_tolerance"] = 1e-4 if mode == "bf16" else 1e-5
        out["loss_rel_err_vs_single_process"] = abs(float(loss_sum) - float(loss1)) / abs(float(loss1))
        out["global_batch_checked"] = int(cat[0].shape[0])
    opt.zero_grad()
    model.train()
    # (3)
    pk.set_compute_mode("fp32")
    n_total = 4 * world
    feats, _ = synthetic.utterances(n_total, np.random.RandomState(97))
    src_all, mask_all = pad_to_longest(feats)                      # one global T: the TDNN encoder depends on the padding
    dopt = types.SimpleNamespace(use_gpu=True, beam_size=10, max_token_seq_len=40, nbest=1)
    lo_, hi_ = parallel.shard_range(n_total, rank, world)
    hyps, _ = translate_batch(model, (None, src_all[lo_:hi_], mask_all[lo_:hi_], None, None), dopt, None)
    parts = [None] * world
    dist.all_gather_object(parts, hyps)
    if rank == 0:
        whole, _ = translate_batch(model, (None, src_all, mask_all, None, None), dopt, None)
        out["sharded_decode_equals_unsharded"] = bool([h for part in parts for h in part] == whole)
        out["decode_utterances_checked"] = n_total
    pk.set_compute_mode(mode)
    if rank == 0:
        out["ok"] = bool(out["params_bit_identical"] and out["sharded_decode_equals_unsharded"]
                         and out["grad_rel_err_vs_single_process"] <= out["grad_tolerance"]
                         and out["loss_rel_err_vs_single_process"] <= 1e-5
                         and out.get("peer_step_max_abs_diff_vs_nccl_sum_plus_adam", 0.0) <= 1e-6)
    return out


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
    torch.manual_seed(0)                                         # identical replicas on every rank
    model = pk.Transformer(lda_mat=synthetic.lda_matrix(), seed=1000 + rank, **model_config()).cuda()
    opt = pk.ScheduledOptim(pk.FusedAdam(model.parameters(), betas=(0.9, 0.999), eps=1e-8), 1e-3, 25000)
    pool = build_pool(B, rank, args.padding)                     # every rank its own utterances
    n_pool = len(pool)
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
    graphed, graph_note, launches_per_step = None, "eager", None
    if not args.no_graph:
        graphed = pk.GraphedTrainStep(model, opt, None, grad_sync=sync.finish if sync else None)
        per_shape = {}
        for b in pool:                                            # one graph per batch shape, captured before timing
            key = graphed._key(b[1], b[3])
            if key not in per_shape:
                n0 = _lib.launch_count()
                graphed.capture(b)
                per_shape[key] = (_lib.launch_count() - n0) // (len(per_shape) and 2 or graphed.warmup + 1)
        launches_per_step = int(np.mean(list(per_shape.values())))
        graph_note = "cuda-graph (%d shape%s)" % (len(per_shape), "s" if len(per_shape) > 1 else "")

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
            graphed.step(*dev_pool[i % n_pool])                   # device->device copy into the static buffers + replay
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

    def timed_block(step_fn, run_fn, first):
        """EXACTLY args.steps steps between barrier + synchronize, CUDA events on the launching stream -> (ms, frames)."""
        import gc
        barrier()
        gc.collect()
        gc.disable()                              # no collector pauses inside the timed region
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        if run_fn is not None:
            run_fn(first, args.steps)
        else:
            for i in range(args.steps):
                step_fn(first + i)
        e.record()
        barrier()
        gc.enable()
        ms = s.elapsed_time(e)
        frames = sum(frames_pool[(first + i) % n_pool] for i in range(args.steps))
        t = torch.tensor([ms, float(frames)], device="cuda", dtype=torch.float64)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms, frames = float(tmax[0]), float(t[1])
        return ms, frames

    def timed(step_fn, run_fn, repeats):
        """Warm-up, then `repeats` timed blocks of exactly args.steps steps each; the reported block is the median one
        (a 20-step block is 20-30 ms: a single block is at the mercy of one host hiccup)."""
        if run_fn is not None:
            run_fn(0, args.warmup)
        else:
            for i in range(args.warmup):
                step_fn(i)
        blocks = [timed_block(step_fn, run_fn, args.warmup + r * args.steps) for r in range(repeats)]
        order = sorted(range(repeats), key=lambda r: blocks[r][0] / blocks[r][1])
        med = blocks[order[repeats // 2]]
        per_step = [b[0] / args.steps for b in blocks]
        return med[0], med[1], {"repeats": repeats, "ms_per_step_min": min(per_step), "ms_per_step_median": float(np.median(per_step)),
                                "ms_per_step_max": max(per_step), "ms_per_step_all": [round(x, 5) for x in per_step]}

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    n0 = _lib.launch_count()
    ms_res, frames_res, blocks_res = timed(resident_step, None, args.repeats)
    eager_launches = (_lib.launch_count() - n0) // max(1, args.repeats + 1)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, frames_e2e, blocks_e2e = timed(None, e2e_run, max(3, args.repeats // 2))

    parity = None
    if world > 1 and not args.no_parity:
        parity = dp_parity_check(model, opt, sync, pool, rank, world, pk, args.mode)

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
    pk.set_compute_mode(args.mode)
    if rank != 0:
        _finish(world, graphed)
        return
    value = frames_res / (ms_res * 1e-3)
    e2e_value = frames_e2e / (ms_e2e * 1e-3)
    b0 = pool[0]
    h2d = float(np.mean([b[1].astype(np.float32).nbytes + b[2].astype(np.uint8).nbytes + b[3].astype(np.int64).nbytes
                         + b[4].astype(np.uint8).nbytes for b in pool]))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.mode == "fp32" else "bf16", "data": "synthetic",
        "config": workload_config(B, world, pool, args.padding),
        "execution": graph_note, "gradient_exchange": "none (1 GPU)" if world == 1 else dp_backend,
        "blocks": blocks_res,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 12,
                "ms_per_step": ms_e2e / args.steps, "blocks": blocks_e2e,
                "api": "train_epoch(model, loader_of_K_pinned_batches, None, 'train', optimizer, graphed=..., sync_every_step=True)"},
        "gpu_launches": int(launches_per_step * args.steps) if launches_per_step else int(eager_launches),
        "launches_per_step": launches_per_step,
        "clocks": clocks,
        "real_frames_per_step_per_gpu": float(np.mean(frames_pool)),
    }
    if parity is not None:
        line["dp_parity"] = bool(parity.get("ok"))
        line["dp_parity_detail"] = parity
    if dec is not None:
        line["decode"] = dec
    if world == 1:
        pk_ = peaks()
        fl = float(np.mean([timit_step_flops(b) for b in pool]))
        tf = fl / (ms_res / args.steps * 1e-3) / 1e12
        line["roofline_step"] = {"bound": "tensor", "achieved": tf, "peak": pk_["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                 "frac": tf / pk_["bf16_tflops_sustained"], "flops_per_step": fl,
                                 "convention": "whole train step, algorithmic FLOPs on real frames / tokens (SURVEY 8d)"}
        if dec is not None:
            try:
                v = cpu_decode_baseline(16, 10, 100, os.cpu_count() or 1)
                dec["cpu_baseline"] = {"value": v, "unit": "utts/s", "cores": os.cpu_count() or 1, "kind": "port",
                                       "sample": "16 utterances in batches of 8, beam 10, 100 forced steps, oracle port of L/decode.py:22-107 (no KV cache)"}
                dec["roofline"] = decode_roofline_probe()
            except Exception as exc:
                dec["cpu_baseline"] = {"error": str(exc)[:200]}
        try:
            Tp = max(int(b[1].shape[1]) for b in pool)
            line["roofline"] = roofline_probe(model, B, Tp, pk, args.mode)
            line["roofline_hbm"] = hbm_probe()
        except Exception as exc:                                   # never lose the headline line to a probe
            line["roofline"] = {"error": str(exc)[:200]}
        if not args.no_cfg5:
            try:
                line["cfg5"] = cfg5_bench(B=8, band=(-100, 0), steps=5, warmup=3)
                line["cfg5_full_attention"] = cfg5_bench(B=4, band=(-1600, 1600), steps=3, warmup=2)
                line["cfg5_attention_kernel"] = {"band_-100_0": attention_probe(band=(-100, 0)), "full": attention_probe(band=None)}
                pk.set_compute_mode(args.mode)
            except Exception as exc:
                line["cfg5"] = {"error": str(exc)[:200]}
        threads = os.cpu_count() or 1
        t0 = time.time()
        v, ms = cpu_train_baseline(pool, 8, 1, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_step": ms,
                                "sample": "8 steps x %d utterances (the same synthetic batches and padding), oracle port of the "
                                          "reference train step, fp32, dropout 0.35, %.0f s of CPU work" % (B, time.time() - t0)}
    print(json.dumps(line))
    _finish(world, graphed)


def _finish(world, graphed=None):
    """Tear down in dependency order: the CUDA graphs hold captured NCCL collectives, so they go first (with the GPU
    idle), then the process group.  PKA_HARD_EXIT=1 restores the old behaviour (flush + os._exit) should a driver ever
    hang in NCCL's destructor."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        import gc
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        if os.environ.get("PKA_HARD_EXIT", "0") == "1":
            os._exit(0)
        if graphed is not None:
            graphed.release()
        gc.collect()
        torch.cuda.synchronize()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--repeats", type=int, default=9, help="timed blocks of --steps steps; the median block is reported")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--mode", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--padding", default=os.environ.get("PKA_BENCH_PADDING", "bucket"), choices=["set", "bucket"],
                    help="bucket: length-sorted batches padded to a few bucket lengths (one CUDA graph each); set: the "
                         "reference loader's whole-set padding")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
