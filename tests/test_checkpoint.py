"""CPU tests of checkpoints and model averaging against fixtures written by the REAL reference
(tests/golden/make_checkpoint_golden.py): a whole-module pickle as L/initialize_model.py saves it, and the running
averages of its own `scale_dict` / `add_dict`."""
import argparse
import os

import numpy as np
import pytest
import torch

from pytorch_kaldi_asr_b200 import checkpoint as ck

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REF_FILE = os.path.join(GOLDEN, "ref_checkpoint_tiny.torch")


def golden_arrays():
    g = np.load(os.path.join(GOLDEN, "checkpoint_tiny.npz"))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    return g, sd


def perturbed(sd, i):          # same recipe as make_checkpoint_golden.py
    out = {}
    for j, (k, v) in enumerate(sd.items()):
        noise = np.random.RandomState(1000 * i + j).randn(*v.shape).astype(np.float32)
        out[k] = v + 0.05 * torch.from_numpy(noise).reshape(v.shape)
    return out


def test_reference_pickled_module_checkpoint_loads_into_this_package():
    g, sd = golden_arrays()
    raw = ck.read_checkpoint(REF_FILE)
    assert raw["format"] == "reference-pickled-module" and raw["epoch"] == 0 and raw["optimizer"] is None
    assert raw["model_options"]["src_dim"] == 4 and raw["model_options"]["tdnn_contexts"] == [[-1, 0, 1], [-3, 0, 3]]
    assert set(raw["state_dict"]) == set(sd)
    for k in sd:
        assert torch.equal(raw["state_dict"][k], sd[k]), k
    kw = ck.model_kwargs(raw["model_options"])
    assert kw["n_src_dim"] == 4 and kw["n_tgt_vocab"] == 9 and kw["encoder_sub_sequence"] == (-100, 0)
    assert kw["decoder_sub_sequence"] == (-3, 0) and kw["en_d_model"] == 16 and "read_vocab_file" not in kw
    np.testing.assert_array_equal(ck.lda_from_state_dict(raw["state_dict"]), g["lda_mat"])
    full = ck.load_checkpoint(REF_FILE)
    model = full["model"]
    assert type(model).__module__ == "pytorch_kaldi_asr_b200.transformer.Models"
    assert hasattr(model, "dropout_state")                       # a properly constructed instance, not the unpickled shell
    got = model.state_dict()
    assert list(got) == list(sd) and all(torch.equal(got[k], sd[k]) for k in sd)


def test_state_dict_checkpoint_round_trip_is_plain_data(tmp_path):
    _, sd = golden_arrays()
    model = ck.load_checkpoint(REF_FILE)["model"]
    path = str(tmp_path / "epoch.3.torch")
    opts = ck.read_checkpoint(REF_FILE)["model_options"]
    ck.save_checkpoint(path, model, argparse.Namespace(**opts), 3, train_options=argparse.Namespace(epoch=50, batch_size=32),
                       extra=dict(note="x"))
    assert os.listdir(str(tmp_path)) == ["epoch.3.torch"]       # the temporary file was renamed into place
    plain = torch.load(path, map_location="cpu", weights_only=True)      # no pickled classes inside
    assert plain["format"] == ck.FORMAT and plain["epoch"] == 3 and plain["train_options"] == dict(epoch=50, batch_size=32)
    assert plain["model_options"]["seed"] == 0 and plain["model_options"]["encoder_type"] == "tdnn"
    again = ck.load_checkpoint(path)
    assert again["extra"] == dict(note="x") and again["optimizer"] is None
    got = again["model"].state_dict()
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    with pytest.raises(ValueError):
        ck.model_kwargs(dict(en_d_model=16))
    torch.save([1, 2, 3], str(tmp_path / "junk.torch"))
    with pytest.raises(ValueError):
        ck.read_checkpoint(str(tmp_path / "junk.torch"))


def test_running_average_equals_the_reference_combine_arithmetic():
    g, sd = golden_arrays()
    seen = 0
    for n, avg in ck.running_average(perturbed(sd, i) for i in range(4)):
        for k in sd:
            want = torch.from_numpy(g["avg%d.%s" % (n, k)])
            assert torch.equal(avg[k], want), (n, k)             # same ops in the same order: bit-exact
        seen = n
    assert seen == 4
    with pytest.raises(ValueError):
        list(ck.running_average([sd, {k: v for k, v in list(sd.items())[:-1]}]))


def test_combine_keeps_the_best_running_average(tmp_path, monkeypatch, capsys):
    """`train.combine` / `combine_files` (L/train.py:284-322, L/combine.py) with the evaluation mocked out: epoch files
    are walked newest first, missing files end the walk, the best (not the last) average is what gets saved."""
    from pytorch_kaldi_asr_b200 import train as T
    _, sd = golden_arrays()
    raw = ck.read_checkpoint(REF_FILE)
    model = ck.build_model(raw["model_options"], sd)
    for epoch in (2, 3, 4, 5):                                   # epoch.1 is missing: the walk from 5 must stop at 2
        ck.save_state(str(tmp_path / ("epoch.%d.torch" % epoch)), perturbed(sd, epoch), model, raw["model_options"], epoch)
    scores = iter([(1.0, 0.20), (0.9, 0.50), (0.8, 0.40), (0.7, 0.50)])
    seen = []

    def fake_epoch(m, data, crit, mode="train", **kw):
        assert mode == "eval" and not m.training is None
        seen.append({k: v.detach().clone() for k, v in m.state_dict().items()})
        return next(scores)
    monkeypatch.setattr(T, "train_epoch", fake_epoch)
    opt = argparse.Namespace(save_model_dir=str(tmp_path))
    best = T.combine(opt, 5, None, data=None, num_model=10, device="cpu")
    assert best == 0.50 and len(seen) == 4
    out = ck.read_checkpoint(str(tmp_path / "combined.accu50.00.torch"))
    assert out["extra"]["averaged_models"] == 2                   # the first 0.50, not the later tie
    assert [os.path.basename(f) for f in out["extra"]["averaged_from"]] == ["epoch.5.torch", "epoch.4.torch"]
    want = list(ck.running_average([perturbed(sd, 5), perturbed(sd, 4)]))[-1][1]
    assert all(torch.equal(out["state_dict"][k], want[k]) for k in want)
    assert all(torch.equal(seen[1][k], want[k]) for k in want)    # the model evaluated second carried that average
    assert out["epoch"] == 5 and out["train_options"] == dict(save_model_dir=str(tmp_path))
    assert "[INFO] averaging 4 models" in capsys.readouterr().out
    with pytest.raises(ValueError):
        T.combine(opt, 9, None, data=None, device="cpu")


def test_train_driver_checkpoints_and_best_snapshot(tmp_path, monkeypatch, capsys):
    """`train.train` (L/train.py:217-272) with the epochs mocked out: which epochs are saved, that the best model is the
    snapshot taken at its epoch (not the live module), that only the writer rank writes, and where a resume starts."""
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200 import train as T
    _, sd = golden_arrays()
    raw = ck.read_checkpoint(REF_FILE)
    model = ck.build_model(raw["model_options"], sd)
    optimizer = pk.ScheduledOptim(torch.optim.Adam(model.parameters(), betas=(0.9, 0.999), eps=1e-8), 1e-3, 100)
    probe = "decoder.tgt_word_emb.weight"
    dev_accuracy = {1: 0.30, 2: 0.35, 3: 0.60, 4: 0.55}
    calls, state = [], dict(epoch=0)

    def fake_epoch(m, data, crit, mode="train", batch_eval=10, **kw):
        calls.append((data, mode, batch_eval))
        if mode == "train":
            state["epoch"] += 1
            with torch.no_grad():
                dict(m.named_parameters())[probe].fill_(float(state["epoch"]))      # "training" marks the weights
            return 1.0, 0.10
        return 1.0, (dev_accuracy[state["epoch"]] if data == "dev" else 0.20)
    monkeypatch.setattr(T, "train_epoch", fake_epoch)
    opt = argparse.Namespace(epoch=4, save_interval=2, save_model_dir=str(tmp_path), seq_error_prob=0)
    best_accu, best_epoch = T.train(model, "train", "dev", "test", None, optimizer, opt, raw["model_options"])
    assert (best_accu, best_epoch) == (0.60, 3)
    assert calls[:4] == [("train", "train", 10), ("train", "eval", 10), ("dev", "eval", 10), ("test", "eval", 10)]
    files = sorted(os.listdir(str(tmp_path)))
    assert files == ["best.epoch3.accu60.00.torch", "epoch.2.torch", "epoch.3.torch", "epoch.4.torch"]
    best = ck.read_checkpoint(str(tmp_path / files[0]))
    assert best["epoch"] == 3 and float(best["state_dict"][probe][0, 0]) == 3.0        # the reference would store 4.0
    assert float(model.state_dict()[probe][0, 0]) == 4.0
    last = ck.read_checkpoint(str(tmp_path / "epoch.4.torch"))
    assert last["optimizer"]["schedule"]["n_current_steps"] == 0 and last["train_options"]["save_interval"] == 2
    assert "best valid accuracy: 60.00 %, on epoch 3" in capsys.readouterr().out

    # a non-writer rank trains and evaluates but writes nothing; a resumed run starts at opt.start_epoch
    for f in files:
        os.remove(str(tmp_path / f))
    state["epoch"], calls[:] = 2, []
    opt.start_epoch = 3
    assert T.train(model, "train", "dev", "test", None, optimizer, opt, raw["model_options"], writer=False) == (0.60, 3)
    assert os.listdir(str(tmp_path)) == [] and sum(1 for c in calls if c[1] == "train") == 2
