"""`-m gpu`: the recipe's command-line entry points end to end on a tiny Kaldi-style data directory (SURVEY.md 8f
ranks 3-4): initialize_model -> train (2 epochs, checkpoints, combine) -> resume -> decode (n-best file) -> rescore ->
score.  Every stage is called through its `main(argv)` with the reference's flag names."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PHONES = ["aa", "b", "ch", "d", "eh", "sil"]


def make_dir(root, name, n_utts, seed, dim=5):
    from pytorch_kaldi_asr_b200.utils import kaldi_ark
    rng = np.random.RandomState(seed)
    d = os.path.join(str(root), name)
    os.makedirs(d)
    items, lines = [], []
    for i in range(n_utts):
        t = int(rng.randint(10, 25))
        items.append(("%s%02d" % (name, i), rng.randn(t, dim).astype(np.float32)))
        lines.append(items[-1][0] + " " + " ".join(PHONES[j] for j in rng.randint(0, len(PHONES), size=t // 3)) + "\n")
    kaldi_ark.write_ark_scp(d + "/feats.ark", d + "/feats.scp", items)
    open(d + "/text", "w").write("".join(lines))
    return d


def test_recipe_end_to_end(tmp_path, capsys):
    from pytorch_kaldi_asr_b200 import checkpoint, results
    from pytorch_kaldi_asr_b200.recipe import decode, initialize_model, rescore, score, train
    from pytorch_kaldi_asr_b200.utils import kaldi_ark
    root = str(tmp_path)
    tr, dev, te = make_dir(root, "train", 16, 1), make_dir(root, "dev", 6, 2), make_dir(root, "test", 6, 3)
    with open(root + "/vocab", "w") as f:
        for i, w in enumerate(["<blank>", "<unk>", "<s>", "</s>"] + PHONES):
            f.write("%s %d\n" % (w, i))
    kaldi_ark.write_mat(root + "/lda.mat", (np.random.RandomState(0).randn(25, 26) * 0.1).astype(np.float32))

    initialize_model.main(["-read_feats_scp_file", tr + "/feats.scp", "-lda_mat_file", root + "/lda.mat",
                           "-read_vocab_file", root + "/vocab", "-encoder_max_len", "40", "-decoder_max_len", "16",
                           "-decoder_sub_sequence", "(-5,0)", "-en_layers", "1", "-de_layers", "1", "-n_head", "2",
                           "-en_d_model", "32", "-de_d_model", "32", "-d_k", "16", "-d_v", "16", "-en_dropout", "0.1",
                           "-de_dropout", "0.1", "-save_model_file", root + "/model.init", "-init_seed", "0"])
    init = checkpoint.read_checkpoint(root + "/model.init")
    assert init["epoch"] == 0 and init["model_options"]["src_dim"] == 5 and init["model_options"]["tgt_vocab_dim"] == 10
    assert len(init["model_options"]["tdnn_contexts"]) == 6

    common = ["-read_train_dir", tr, "-read_dev_dir", dev, "-read_test_dir", te, "-read_vocab_file", root + "/vocab",
              "-save_model_dir", root + "/exp", "-epoch", "2", "-optim_start_lr", "0.002", "-batch_size", "4",
              "-save_interval", "1", "-use_gpu"]
    best = train.main(common + ["-load_model_file", root + "/model.init"])
    files = sorted(os.listdir(root + "/exp"))
    assert "epoch.1.torch" in files and "epoch.2.torch" in files
    assert sum(f.startswith("best.") for f in files) == 1 and sum(f.startswith("combined.") for f in files) == 1
    assert 0.0 <= best <= 1.0
    first = checkpoint.read_checkpoint(root + "/exp/epoch.1.torch")
    assert first["optimizer"]["schedule"]["n_current_steps"] == 4              # 16 utterances / batch 4

    # resume from the epoch-1 checkpoint: epoch 2 runs again from the saved optimiser / schedule / dropout state
    before = checkpoint.read_checkpoint(root + "/exp/epoch.2.torch")
    for f in files:
        if f.startswith("combined.") or f.startswith("best."):
            os.remove(root + "/exp/" + f)
    train.main(common + ["-load_model_file", root + "/exp/epoch.1.torch", "-resume"])
    after = checkpoint.read_checkpoint(root + "/exp/epoch.2.torch")
    assert after["optimizer"]["schedule"]["n_current_steps"] == 8 == before["optimizer"]["schedule"]["n_current_steps"]
    combined = [f for f in os.listdir(root + "/exp") if f.startswith("combined.")]
    assert len(combined) == 1

    n = decode.main(["-read_data_dir", te, "-read_vocab_file", root + "/vocab", "-load_model_file",
                     root + "/exp/" + combined[0], "-save_result_file", root + "/exp/decode.txt", "-max_token_seq_len", "12",
                     "-batch_size", "4", "-beam_size", "4", "-nbest", "2", "-use_gpu"])
    assert n == 6
    table = results.read_nbest(root + "/exp/decode.txt")
    assert sorted(table) == ["test%02d" % i for i in range(6)]
    for scores, hyps in table.values():
        assert 1 <= len(hyps) <= 2 and all(np.isfinite(scores))
        assert all(w in PHONES or w == "<unk>" for h in hyps for w in h.split())
    n_lines = sum(len(h) for _, h in table.values())
    open(root + "/exp/lm.txt", "w").write("0.0\n" * n_lines)
    os.makedirs(root + "/exp/scoring")
    out = rescore.main(["-decode_file", root + "/exp/decode.txt", "-lm_score", root + "/exp/lm.txt", "-save_dir",
                        root + "/exp/scoring", "-inv_weight_list", "10,20"])
    assert len(out) == 2
    top1 = {k: h[int(np.argmax(s))] for k, (s, h) in table.items()}          # a zero LM keeps the model's best hypothesis
    assert dict(line.rstrip("\n").split(" ", 1) if " " in line.rstrip("\n") else (line.strip(), "")
                for line in open(out[0])) == top1
    name, wer, line = score.main(["-text", te + "/text", "-scoring_dir", root + "/exp/scoring", "-result_file",
                                  root + "/exp/result.txt"])
    assert 0.0 <= wer and line.startswith("%WER") and os.path.exists(root + "/exp/scoring/rescore_10.0_wer")
    assert open(root + "/exp/result.txt").read().startswith("[INFO] best wer presented in file:")
    assert "[PROCEDURE] combining start on best epoch" in capsys.readouterr().out


def test_stage_driver_runs_init_train_decode_rescore_score_with_the_reference_layout(tmp_path, capsys):
    """recipe.run = stages 3-5 of P/run.sh:64-203: exp/model_*/model.init -> epoch.N / combined.* -> decode_{dev,test}/
    decode.txt, lm.3k.score.txt, scoring/rescore_<w>[_wer], result.txt; a toy LM command exercises the score pipe."""
    from pytorch_kaldi_asr_b200.recipe import run
    from pytorch_kaldi_asr_b200.utils import kaldi_ark
    root = str(tmp_path)
    os.makedirs(root + "/data/lang")
    for name, n, seed in (("train_filtered", 16, 1), ("dev_filtered", 6, 2), ("test_filtered", 6, 3)):
        make_dir(root + "/data", name, n, seed)
    with open(root + "/data/lang/vocab.txt", "w") as f:
        for i, w in enumerate(["<blank>", "<unk>", "<s>", "</s>"] + PHONES):
            f.write("%s %d\n" % (w, i))
    kaldi_ark.write_mat(root + "/data/lda.mat", (np.random.RandomState(0).randn(25, 26) * 0.1).astype(np.float32))
    summary = run.main(["-data_dir", root + "/data", "-lang_dir", root + "/data/lang", "-exp_dir", root + "/exp",
                        "-model_suffix", "_tiny", "-encoder_max_len", "40", "-decoder_max_len", "16", "-decoder_sub_sequence",
                        "(-5,0)", "-en_layers", "1", "-de_layers", "1", "-en_d_model", "32", "-de_d_model", "32", "-d_k", "16",
                        "-d_v", "16", "-en_dropout", "0.1", "-de_dropout", "0.1", "-init_seed", "0", "-epoch", "2",
                        "-batch_size", "4", "-optim_start_lr", "0.002", "-compute_mode", "fp32", "-max_token_seq_len", "12",
                        "-decode_batch_size", "4", "-beam_size", "4", "-nbest", "2", "-inv_weight_list", "10,20",
                        "-lm_score_cmd", "awk '{print -NF}'"])
    model_dir = summary["model_dir"]
    assert os.path.basename(model_dir).startswith("model_") and model_dir.endswith("_tiny")
    files = os.listdir(model_dir)
    assert "model.init" in files and "epoch.2.torch" in files and sum(f.startswith("combined.") for f in files) == 1
    for name in ("dev", "test"):
        d = model_dir + "/decode_" + name
        n_hyp = sum(1 for _ in open(d + "/decode.txt"))
        assert n_hyp >= 6 and sum(1 for _ in open(d + "/lm.3k.score.txt")) == n_hyp
        assert sorted(f for f in os.listdir(d + "/scoring")) == ["rescore_10.0", "rescore_10.0_wer", "rescore_20.0", "rescore_20.0_wer"]
        assert open(d + "/result.txt").read().startswith("[INFO] best wer presented in file:")
        assert 0.0 <= summary["wer_" + name]
    out = capsys.readouterr().out
    assert "[PROCEDURE] decoding dev set" in out and "[INFO] language model score computed." in out
