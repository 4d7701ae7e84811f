"""`-m gpu`: checkpoint / resume / model averaging on the device.

* a run resumed from a checkpoint (weights + Adam moments + LR-schedule step + dropout counter) continues bit-exactly;
* FusedAdam's state dict is torch.optim.Adam's: it loads into a stock Adam and back, and the next step agrees;
* the `train` / `combine` drivers (L/train.py:217-322) write loadable files and `combine` keeps the best running average.
"""
import argparse
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REF_FILE = os.path.join(GOLDEN, "ref_checkpoint_tiny.torch")
BETAS, EPS = (0.9, 0.999), 1e-8


class ListLoader(list):
    mode = "drop"


def tiny_batches(n, batch=4, seed=3):
    from pytorch_kaldi_asr_b200.utils import synthetic
    return synthetic.batches(n, batch, seed=seed, pad_to="set", feat_dim=4, vocab=9, mean_len=14.0, std_len=4.0,
                             min_len=8, max_len=20, label_div=4, max_labels=8)


def fresh(options=None, state_dict=None):
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200 import checkpoint as ck
    raw = ck.read_checkpoint(REF_FILE)
    opts = dict(raw["model_options"], **(options or {}))
    model = ck.build_model(opts, state_dict or raw["state_dict"], device="cuda")
    return model, opts


def weights(model):
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def test_resume_is_bit_exact(tmp_path):
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200 import checkpoint as ck
    batches = tiny_batches(6)

    def optimizer_for(model):
        return pk.ScheduledOptim(pk.FusedAdam(model.parameters(), betas=BETAS, eps=EPS), 2e-3, 10)

    def steps(model, opt, some):
        return [pk.train_epoch(model, ListLoader([b]), None, mode="train", optimizer=opt)[0] for b in some]

    model, opts = fresh()
    assert model.decoder is not None and opts["en_dropout"] == 0.1            # dropout is on: the counter matters
    opt = optimizer_for(model)
    straight = steps(model, opt, batches)
    w_straight = weights(model)

    model, opts = fresh()
    opt = optimizer_for(model)
    first = steps(model, opt, batches[:3])
    path = str(tmp_path / "epoch.1.torch")
    ck.save_checkpoint(path, model, opts, 1, optimizer=opt)
    del model, opt

    loaded = ck.load_checkpoint(path, device="cuda")
    assert loaded["optimizer"]["schedule"]["n_current_steps"] == 3
    assert loaded["optimizer"]["adam"]["fused"]["adam_t"] == 3
    model2 = loaded["model"]
    opt2 = optimizer_for(model2)
    ck.restore_optimizer(opt2, loaded["optimizer"], model2)
    assert opt2.n_current_steps == 3 and opt2.optimizer.param_groups[0]["lr"] == pytest.approx(2e-3 * 10 / 13)
    second = steps(model2, opt2, batches[3:])
    assert first + second == straight, (first, second, straight)
    w_resumed = weights(model2)
    for k in w_straight:
        assert torch.equal(w_straight[k], w_resumed[k]), k


def test_fused_adam_state_interchanges_with_torch_adam():
    import pytorch_kaldi_asr_b200 as pk
    batches = tiny_batches(3, seed=5)
    no_drop = dict(en_dropout=0.0, de_dropout=0.0)

    def run(model, opt, some):
        for b in some:
            pk.train_epoch(model, ListLoader([b]), None, mode="train", optimizer=opt)

    class Plain:                                                 # train_epoch's optimizer protocol around a bare optimizer
        def __init__(self, inner):
            self.optimizer = inner

        def zero_grad(self):
            self.optimizer.zero_grad()

        def step(self):
            self.optimizer.step()

        def update_learning_rate(self):
            pass

    for direction in ("fused->torch", "torch->fused"):
        a, opts = fresh(no_drop)
        make_a = (lambda p: pk.FusedAdam(p, lr=1e-3, betas=BETAS, eps=EPS)) if direction == "fused->torch" else \
            (lambda p: torch.optim.Adam(p, lr=1e-3, betas=BETAS, eps=EPS))
        make_b = (lambda p: torch.optim.Adam(p, lr=1e-3, betas=BETAS, eps=EPS)) if direction == "fused->torch" else \
            (lambda p: pk.FusedAdam(p, lr=1e-3, betas=BETAS, eps=EPS))
        opt_a = make_a(a.parameters())
        run(a, Plain(opt_a), batches[:2])
        state = opt_a.state_dict()
        if direction == "fused->torch":                          # one entry per trainable parameter, torch's indexing
            index = {id(p): i for i, p in enumerate(a.parameters())}
            assert sorted(state["state"]) == sorted(index[id(p)] for p in a.parameters() if p.requires_grad)
        b, _ = fresh(no_drop, weights(a))
        opt_b = make_b(b.parameters())
        opt_b.load_state_dict({k: v for k, v in state.items() if k != "fused"})
        run(a, Plain(opt_a), batches[2:])
        run(b, Plain(opt_b), batches[2:])
        wa, wb = weights(a), weights(b)
        for k in wa:
            assert torch.allclose(wa[k], wb[k], rtol=1e-5, atol=1e-7), (direction, k)
        back = opt_b.state_dict()["state"]
        assert all(float(v["step"]) == 3.0 for v in back.values())


def test_train_and_combine_drivers(tmp_path, capsys):
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200 import checkpoint as ck, train as tr
    model, opts = fresh()
    optimizer = pk.ScheduledOptim(pk.FusedAdam(model.parameters(), betas=BETAS, eps=EPS), 5e-3, 100)
    train_data, dev_data, test_data = (ListLoader(tiny_batches(n, seed=s)) for n, s in ((5, 1), (2, 2), (2, 3)))
    opt = argparse.Namespace(epoch=3, save_interval=1, save_model_dir=str(tmp_path), seq_error_prob=0, use_gpu=True)
    best_accu, best_epoch = tr.train(model, train_data, dev_data, test_data, None, optimizer, opt, opts)
    files = sorted(os.listdir(str(tmp_path)))
    assert [f for f in files if f.startswith("epoch.")] == ["epoch.1.torch", "epoch.2.torch", "epoch.3.torch"]
    best_file = [f for f in files if f.startswith("best.")]
    assert best_file == ["best.epoch%d.accu%.2f.torch" % (best_epoch, 100 * best_accu)]
    best = ck.load_checkpoint(str(tmp_path / best_file[0]), device="cuda")
    snap = ck.read_checkpoint(str(tmp_path / ("epoch.%d.torch" % best_epoch)))["state_dict"]
    assert all(torch.equal(best["state_dict"][k], snap[k]) for k in snap)      # the best epoch's weights, not the last's
    _, accu = pk.train_epoch(best["model"], dev_data, None, mode="eval")
    assert accu == pytest.approx(best_accu, abs=1e-12)

    comb_accu = tr.combine(opt, 3, None, dev_data, num_model=3)
    comb_file = [f for f in os.listdir(str(tmp_path)) if f.startswith("combined.")]
    assert comb_file == ["combined.accu%.2f.torch" % (100 * comb_accu)]
    comb = ck.load_checkpoint(str(tmp_path / comb_file[0]), device="cuda")
    n = comb["extra"]["averaged_models"]
    assert 1 <= n <= 3 and [os.path.basename(f) for f in comb["extra"]["averaged_from"]] == \
        ["epoch.%d.torch" % e for e in (3, 2, 1)[:n]]
    sds = [{k: v.cuda() for k, v in ck.read_checkpoint(str(tmp_path / ("epoch.%d.torch" % e)))["state_dict"].items()}
           for e in (3, 2, 1)]
    accus = []
    for m, avg in ck.running_average(sds):                       # on the device, like combine: the same arithmetic
        probe, _ = fresh(state_dict=avg)
        accus.append(pk.train_epoch(probe, dev_data, None, mode="eval")[1])
        if m == n:
            assert all(torch.equal(comb["state_dict"][k], avg[k].cpu()) for k in avg)
    assert comb_accu == pytest.approx(max(accus), abs=1e-12) and accus.index(max(accus)) + 1 == n
    out = capsys.readouterr().out
    assert "[INFO] averaging 3 models" in out and "best valid accuracy" in out
