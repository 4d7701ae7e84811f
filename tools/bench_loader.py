#!/usr/bin/env python
"""Host-side throughput of the input feed on a TIMIT-shaped synthetic set, CPU only:

    python tools/bench_loader.py [--utts 3696] [--batch 32] [--epochs 3] [--reference]

Writes a seeded Kaldi archive under a temporary directory, then times (a) the ark/scp reader, (b) this package's
BatchLoader: start-up (pre-load) and steady-state batches per epoch for whole-set and per-batch padding, and with
`--reference` (build container only, needs /root/reference) (c) the reference's BatchLoader on the same files with its
`kaldi_io.read_mat` served by the same reader, so the difference is the loader logic alone.  Prints one JSON object.
"""
import argparse
import json
import os
import random
import sys
import tempfile
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def epoch_rate(loader, epochs, real_frames):
    first_batch, per_epoch = [], []
    for _ in range(epochs):
        t0 = time.perf_counter()
        it = iter(loader)
        n = 0
        for batch in it:
            if n == 0:
                first_batch.append(time.perf_counter() - t0)
            n += 1
        per_epoch.append(time.perf_counter() - t0)
    best = min(per_epoch)
    return dict(batches=n, epoch_s=round(best, 4), ms_per_batch=round(1e3 * best / max(n, 1), 3),
                first_batch_ms=round(1e3 * min(first_batch), 3), real_frames_per_s=round(real_frames / best))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=3696)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--reference", action="store_true")
    args = ap.parse_args()

    from pytorch_kaldi_asr_b200.utils import kaldi_ark, synthetic
    from pytorch_kaldi_asr_b200.utils.BatchLoader import BatchLoader

    rng = np.random.RandomState(1234)
    feats, labels = synthetic.utterances(args.utts, rng)
    keys = ["utt%05d" % i for i in range(args.utts)]
    out = dict(utts=args.utts, batch=args.batch, frames=int(sum(f.shape[0] for f in feats)), cores=os.cpu_count())
    with tempfile.TemporaryDirectory() as tmp:
        ark, scp = tmp + "/feats.ark", tmp + "/feats.scp"
        kaldi_ark.write_ark_scp(ark, scp, zip(keys, feats))
        out["ark_mb"] = round(os.path.getsize(ark) / 2 ** 20, 1)
        table = kaldi_ark.read_scp(scp)
        order = list(table.values())
        random.Random(0).shuffle(order)
        t0 = time.perf_counter()
        n_frames = sum(kaldi_ark.read_mat(rx).shape[0] for rx in order)
        dt = time.perf_counter() - t0
        out["reader"] = dict(seconds=round(dt, 4), frames_per_s=round(n_frames / dt), mb_per_s=round(out["ark_mb"] / dt))
        triples = [(k, table[k], lab) for k, lab in zip(keys, labels)]
        n_used = args.utts // args.batch * args.batch

        t0 = time.perf_counter()
        ours = BatchLoader(triples, args.batch, print_info=False)
        out["ours_preload_s"] = round(time.perf_counter() - t0, 3)
        out["ours_dataset_pad"] = epoch_rate(ours, args.epochs, out["frames"] * n_used / args.utts)
        per_batch = BatchLoader(triples, args.batch, print_info=False, pad_to="batch", bucket=8)
        out["ours_batch_pad_bucket8"] = epoch_rate(per_batch, args.epochs, out["frames"] * n_used / args.utts)
        stream = BatchLoader(triples, args.batch, pre_load=False, print_info=False, read_ahead=4)
        out["ours_streaming_read_ahead4"] = epoch_rate(stream, args.epochs, out["frames"] * n_used / args.utts)

        if args.reference:
            R = os.environ.get("PKA_REFERENCE", "/root/reference")
            stub = types.ModuleType("kaldi_io")
            stub.read_mat = kaldi_ark.read_mat
            sys.modules["kaldi_io"] = stub
            for name in [m for m in sys.modules if m == "utils" or m.startswith("utils.")]:
                del sys.modules[name]
            sys.path.insert(0, R + "/pytorch")
            from utils.BatchLoader import BatchLoader as RefLoader
            t0 = time.perf_counter()
            ref = RefLoader(triples, args.batch, pre_load=True, print_info=False)
            out["reference_preload_s"] = round(time.perf_counter() - t0, 3)
            out["reference_dataset_pad"] = epoch_rate(ref, args.epochs, out["frames"] * n_used / args.utts)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
