#!/usr/bin/env python
"""Generate tests/golden/loader_small.npz by running the REAL reference loader (U/BatchLoader.py, the label helpers of
U/instances_handler.py and `initialize_batch_loader` of L/train.py) in place on a tiny seeded data directory.

Run in the build container only (needs /root/reference):

    python tests/golden/make_loader_golden.py

The reference reads features through the external `kaldi_io` package, which is not installed here; the stub module it
gets is backed by this repo's reader (`utils/kaldi_ark.read_mat`), so the fixture pins the LOADER logic (label
preparation, matching, padding, shuffling, batching, drop/all) -- not the file format.  The data directory is rebuilt
from the seed by `tests/test_loader.py::make_data_dir`, which this script imports, so nothing but the batches is stored.
"""
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
R = os.environ.get("PKA_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import tempfile
    from pytorch_kaldi_asr_b200.utils import kaldi_ark
    from test_loader import make_data_dir, GOLDEN_CASES

    stub = types.ModuleType("kaldi_io")
    stub.read_mat = kaldi_ark.read_mat
    sys.modules["kaldi_io"] = stub
    for name in [m for m in sys.modules if m == "utils" or m.startswith("utils.")]:
        del sys.modules[name]
    sys.path[:0] = [R + "/pytorch", R + "/project/attention-transformer-timit/local/pytorch",
                    R + "/project/attention-transformer-timit/local"]
    import train as ref_train                                  # reference L/train.py
    from utils.BatchLoader import BatchLoader as RefLoader     # reference U/BatchLoader.py

    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        d = make_data_dir(tmp)
        for name, case in GOLDEN_CASES.items():
            random.seed(case["seed"])
            loader = ref_train.initialize_batch_loader(d + "/feats.scp", d + "/text", d + "/vocab", case["batch_size"],
                                                       mode=case["mode"])
            if not case["pre_load"]:
                # the reference hard-wires pre_load=True in initialize_batch_loader; rebuild the streaming variant from
                # the same triples through its own constructor
                keys = list(loader.data["key"])
                scp = kaldi_ark.read_scp(d + "/feats.scp")
                labels = {k: np.asarray(t)[np.asarray(m) > 0] for k, t, m in
                          zip(keys, loader.data["tgt_seq"], loader.data["tgt_pad_mask"])}
                triples = [(k, scp[k], labels[k]) for k in keys]
                random.seed(case["seed"])
                loader = RefLoader(triples, case["batch_size"], pre_load=False, print_info=False, mode=case["mode"])
            for epoch in range(case["epochs"]):
                for n, (key, src, smask, tgt, tmask) in enumerate(loader):
                    p = "%s/e%d/b%d/" % (name, epoch, n)
                    out[p + "key"] = np.array(list(key))
                    out[p + "src"] = np.asarray(src, dtype=np.float32)
                    out[p + "src_mask"] = np.asarray(smask, dtype=np.uint8)
                    out[p + "tgt"] = np.asarray(tgt, dtype=np.int64)
                    out[p + "tgt_mask"] = np.asarray(tmask, dtype=np.uint8)
                out["%s/e%d/n" % (name, epoch)] = np.array(n + 1)
    np.savez_compressed(os.path.join(HERE, "loader_small.npz"), **out)
    print("wrote loader_small.npz with %d arrays" % len(out))


if __name__ == "__main__":
    main()
