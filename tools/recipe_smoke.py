#!/usr/bin/env python
"""Smoke run of the recipe entry points on N GPUs of one box:

    python tools/recipe_smoke.py [--gpus 2] [--out gpurun_out/recipe_n2.json]

Builds a tiny Kaldi-style data directory, initialises a model, then launches `recipe.train` and `recipe.decode` under
torchrun exactly as a user would, and checks what they wrote.  Every child runs under a timeout."""
import argparse
import json
import os
import signal
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def launch(n_gpus, module, args, timeout, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n_gpus), "--master-addr",
           "127.0.0.1", "--master-port", str(port), "-m", module] + args
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    t0 = time.time()
    proc = subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, start_new_session=True)
    try:
        out, _ = proc.communicate(timeout=timeout)
        rc = proc.returncode
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)                  # the launcher and every rank it started (own session)
        out, _ = proc.communicate()
        rc = -9
    return rc, time.time() - t0, out[-3000:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--out", default=None)
    ap.add_argument("--timeout", type=int, default=50)
    args = ap.parse_args()
    from test_gpu_zz_recipe import PHONES, make_dir
    from pytorch_kaldi_asr_b200 import checkpoint, results
    from pytorch_kaldi_asr_b200.recipe import initialize_model
    from pytorch_kaldi_asr_b200.utils import kaldi_ark
    report = dict(gpus=args.gpus)
    with tempfile.TemporaryDirectory() as root:
        tr, dev, te = make_dir(root, "train", 32, 1), make_dir(root, "dev", 6, 2), make_dir(root, "test", 10, 3)
        with open(root + "/vocab", "w") as f:
            for i, w in enumerate(["<blank>", "<unk>", "<s>", "</s>"] + PHONES):
                f.write("%s %d\n" % (w, i))
        kaldi_ark.write_mat(root + "/lda.mat", (np.random.RandomState(0).randn(25, 26) * 0.1).astype(np.float32))
        initialize_model.main(["-read_feats_scp_file", tr + "/feats.scp", "-lda_mat_file", root + "/lda.mat",
                               "-read_vocab_file", root + "/vocab", "-encoder_max_len", "40", "-decoder_max_len", "16",
                               "-decoder_sub_sequence", "(-5,0)", "-en_layers", "1", "-de_layers", "1", "-n_head", "2",
                               "-en_d_model", "32", "-de_d_model", "32", "-d_k", "16", "-d_v", "16", "-save_model_file",
                               root + "/model.init", "-init_seed", "0"])
        rc, secs, tail = launch(args.gpus, "pytorch_kaldi_asr_b200.recipe.train",
                                ["-read_train_dir", tr, "-read_dev_dir", dev, "-read_test_dir", te, "-read_vocab_file",
                                 root + "/vocab", "-load_model_file", root + "/model.init", "-save_model_dir", root + "/exp",
                                 "-epoch", "2", "-batch_size", "4", "-save_interval", "1", "-use_gpu", "-shuffle_seed", "3"],
                                args.timeout, 29611)
        report["train"] = dict(rc=rc, seconds=round(secs, 1), tail=tail if rc else tail[-400:])
        files = sorted(os.listdir(root + "/exp")) if os.path.isdir(root + "/exp") else []
        report["files"] = files
        combined = [f for f in files if f.startswith("combined.")]
        if rc == 0 and combined:
            steps = checkpoint.read_checkpoint(root + "/exp/epoch.2.torch")["optimizer"]["schedule"]["n_current_steps"]
            report["steps_after_2_epochs"] = steps                  # 32 utterances / (batch 4 x N ranks) per epoch
            rc, secs, tail = launch(args.gpus, "pytorch_kaldi_asr_b200.recipe.decode",
                                    ["-read_data_dir", te, "-read_vocab_file", root + "/vocab", "-load_model_file",
                                     root + "/exp/" + combined[0], "-save_result_file", root + "/exp/decode.txt",
                                     "-max_token_seq_len", "12", "-batch_size", "2", "-beam_size", "4", "-nbest", "2",
                                     "-use_gpu"], args.timeout, 29612)
            report["decode"] = dict(rc=rc, seconds=round(secs, 1), tail=tail if rc else tail[-300:])
            if rc == 0:
                table = results.read_nbest(root + "/exp/decode.txt")
                report["decoded_keys"] = len(table)
                report["leftover_parts"] = [f for f in os.listdir(root + "/exp") if f.startswith("decode.txt.")]
    ok = (report["train"]["rc"] == 0 and report.get("decode", {}).get("rc") == 0 and report.get("decoded_keys") == 10
          and report.get("steps_after_2_epochs") == 2 * (32 // (4 * args.gpus)) and not report.get("leftover_parts"))
    report["ok"] = bool(ok)
    text = json.dumps(report, indent=1)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        open(args.out, "w").write(text)
    print(text)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
