"""Whole-path parity on the B200 (`-m gpu`): Transformer forward / loss / backward, optimiser steps and the CUDA-graph
step against (1) the golden fixtures produced by the REAL reference and (2) the CPU oracle at the full TIMIT config.

Tolerances (fp32 path): logits rtol 1e-4 of the logit scale, loss 2e-5 relative, gradients 1e-3 relative to the
per-tensor max-abs, integer outputs exact.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import golden_state_dict, load_golden          # noqa: E402
from oracle import acoustic_model as am                      # noqa: E402
from oracle import train_step as otrain                      # noqa: E402

DEV = "cuda"
SMALL = dict(n_src_dim=8, n_tgt_vocab=11, encoder_max_len=40, decoder_max_len=24, src_fold=1,
             encoder_sub_sequence=(-100, 0), decoder_sub_sequence=(-3, 0), en_layers=2, de_layers=2, n_head=2,
             en_d_model=32, de_d_model=32, d_k=16, d_v=16, en_dropout=0.0, de_dropout=0.0,
             tdnn_contexts=[[-1, 0, 1], [-3, 0, 3]])
FOLD2 = dict(SMALL, src_fold=2, encoder_max_len=20, decoder_sub_sequence=(-2, 0))


def rel_err(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-6))


def build(cfg, sd, lda_mat, **kw):
    import pytorch_kaldi_asr_b200 as pk
    model = pk.Transformer(lda_mat=lda_mat, **cfg, **kw)
    missing = model.load_state_dict(sd, strict=True)
    return model.to(DEV)


def batch_to_dev(g):
    src = torch.from_numpy(g["src"]).to(DEV)
    smask = torch.from_numpy(g["src_mask"]).to(DEV)
    tgt = torch.from_numpy(g["tgt"]).to(DEV)
    tmask = torch.from_numpy(g["tgt_mask"]).to(DEV)
    return src, smask, tgt, tmask


@pytest.mark.parametrize("name,cfg", [("tdnn_small_fwd_bwd", SMALL), ("tdnn_small_fold2_fwd_bwd", FOLD2)])
@pytest.mark.parametrize("smoothing", [False, True])
def test_transformer_vs_reference_golden(name, cfg, smoothing):
    import pytorch_kaldi_asr_b200 as pk
    g = load_golden(name)
    model = build(cfg, golden_state_dict(g), g["lda_mat"])
    model.eval()
    src, smask, tgt, tmask = batch_to_dev(g)
    pred = model(src, smask, tgt[:, :-1], tmask[:, :-1])
    assert rel_err(pred, g["logits"]) <= 1e-4
    loss, n_correct = pk.get_performance(None, pred, tgt[:, 1:], smoothing=smoothing)
    loss.backward()
    tag = "smooth." if smoothing else "plain."
    assert abs(float(loss) - float(g[tag + "loss"])) <= 2e-5 * abs(float(g[tag + "loss"]))
    assert int(n_correct) == int(g[tag + "n_correct"])
    worst = 0.0
    for k, p in model.named_parameters():
        key = tag + "grad." + k
        if key in g.files:
            assert p.grad is not None, k
            e = rel_err(p.grad, g[key])
            worst = max(worst, e)
            assert e <= 1e-3, "%s grad rel err %.2e" % (k, e)
        else:
            assert p.grad is None, "%s is frozen in the reference" % k


def test_attention_encoder_vs_reference_golden():
    from pytorch_kaldi_asr_b200.transformer.Models import Decoder, Encoder
    from pytorch_kaldi_asr_b200 import ops
    g = load_golden("attn_encoder_small")
    enc = Encoder(n_src_dim=8, encoder_max_len=40, n_layers=2, n_head=2, sub_sequence=(-4, 1), d_k=16, d_v=16,
                  d_model=32, d_inner_hid=32, dropout=0.0)
    dec = Decoder(n_tgt_vocab=11, decoder_max_len=24, n_layers=1, n_head=2, sub_sequence=(-3, 0), d_k=16, d_v=16,
                  en_d_model=32, de_d_model=32, d_inner_hid=32, dropout=0.0)
    enc.load_state_dict(golden_state_dict(g, "sd.encoder."))
    dec.load_state_dict(golden_state_dict(g, "sd.decoder."))
    enc, dec = enc.to(DEV).eval(), dec.to(DEV).eval()
    src, smask, tgt, tmask = batch_to_dev(g)
    enc_out, = enc(src, smask)
    assert rel_err(enc_out, g["enc_out"]) <= 1e-4
    logits, = dec(tgt[:, :-1], tmask[:, :-1], smask, enc_out)
    assert rel_err(logits, g["logits"]) <= 1e-4
    loss, _ = ops.cross_entropy_sum(logits.view(-1, logits.size(-1)), tgt[:, 1:].contiguous().view(-1), False)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 2e-5 * abs(float(g["loss"]))
    for mod, pre in ((enc, "grad.encoder."), (dec, "grad.decoder.")):
        for k, p in mod.named_parameters():
            if pre + k in g.files:
                assert rel_err(p.grad, g[pre + k]) <= 1e-3, k


def test_train_steps_vs_reference_golden():
    """4 optimiser steps (Adam + reference LR schedule) reproduce the reference's per-step loss/accuracy."""
    import pytorch_kaldi_asr_b200 as pk
    g = load_golden("train_steps_small")
    model = build(SMALL, golden_state_dict(g, "sd0."), g["lda_mat"])
    opt = pk.ScheduledOptim(pk.FusedAdam(model.parameters(), betas=(0.9, 0.999), eps=1e-8), 2e-3, 10)

    class Loader(list):
        mode = "drop"

    for i in range(4):
        b = (None,) + tuple(g["batch%d.%s" % (i, k)] for k in "src src_mask tgt tgt_mask".split())
        lpw, acc = pk.train_epoch(model, Loader([b]), None, mode="train", optimizer=opt)
        assert abs(lpw - g["per_step"][i, 0]) <= 5e-4 * abs(g["per_step"][i, 0]), (i, lpw, g["per_step"][i, 0])
        assert abs(acc - g["per_step"][i, 1]) <= 1e-9 + 0.02, (i, acc, g["per_step"][i, 1])
    ref_after = golden_state_dict(g, "sd4.")
    sd = model.state_dict()
    for k in am.trainable_keys(ref_after):
        upd = float((ref_after[k] - torch.from_numpy(g["sd0." + k])).abs().max())
        diff = (sd[k].cpu() - ref_after[k]).abs()
        assert float(diff.mean()) <= 2e-3 * upd + 1e-7, k
        assert float(diff.max()) <= 0.25 * upd + 1e-6, k
    batches = [(None,) + tuple(g["batch%d.%s" % (i, k)] for k in "src src_mask tgt tgt_mask".split()) for i in range(4)]
    ev = pk.train_epoch(model, Loader(batches), None, mode="eval", batch_eval=3)
    assert abs(ev[0] - g["eval_after"][0]) <= 2e-3 * abs(g["eval_after"][0])


def timit_setup(B=6, seed=0, dropout=0.0):
    from pytorch_kaldi_asr_b200.utils import synthetic
    cfg = am.example_config(en_dropout=dropout, de_dropout=dropout)
    lda = synthetic.lda_matrix(40, 1, 0)
    sd = am.init_state_dict(cfg, lda, seed=seed)
    batch = synthetic.batches(1, B, seed=1234)[0]
    return cfg, lda, sd, batch


def model_kwargs(cfg):
    return {k: v for k, v in cfg.items() if k != "encoder_type"}


def test_timit_config_logits_loss_grads_vs_oracle():
    """Full example model (P/run.sh:77-91) on TIMIT-shaped synthetic utterances, eval mode."""
    import pytorch_kaldi_asr_b200 as pk
    cfg, lda, sd, batch = timit_setup()
    logits_ref, loss_ref, nc_ref, nw_ref, grads_ref = otrain.loss_and_grads(sd, cfg, batch[1:], smoothing=False)
    model = build(model_kwargs(cfg), sd, lda)
    model.eval()
    src, smask, tgt, tmask = pk.train._to_device(batch, DEV)
    pred = model(src, smask, tgt[:, :-1], tmask[:, :-1])
    assert rel_err(pred, logits_ref) <= 1e-4
    loss, n_correct = pk.get_performance(None, pred, tgt[:, 1:], smoothing=False)
    loss.backward()
    assert abs(float(loss) - float(loss_ref)) <= 2e-5 * abs(float(loss_ref))
    assert int(n_correct) == nc_ref
    for k, p in model.named_parameters():
        if k in grads_ref:
            assert rel_err(p.grad, grads_ref[k]) <= 1e-3, k


def test_train_mode_dropout_parity_with_injected_masks():
    """Dropout ON (p=0.35, all 25 sites): the Philox keep-bits the kernels use are materialised and injected into the
    oracle, so logits/loss/gradients must agree like in eval mode."""
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200 import ops
    cfg, lda, sd, batch = timit_setup(B=3, dropout=0.35)
    model = build(model_kwargs(cfg), sd, lda, seed=17)
    model.train()
    src, smask, tgt, tmask = pk.train._to_device(batch, DEV)
    pred = model(src, smask, tgt[:, :-1], tmask[:, :-1])
    loss, _ = pk.get_performance(None, pred, tgt[:, 1:], smoothing=False)
    loss.backward()
    B, T, L = src.shape[0], src.shape[1], tgt.shape[1] - 1
    H = cfg["n_head"]
    shapes = {"enc.src": (B, T, 256), "enc.out": (B, T, 256), "dec.emb": (B, L, 128), "dec.out": (B, L, 128)}
    for i in range(6):
        shapes["enc.tdnn.%d" % i] = (B, T, 256)
    for l in range(3):
        shapes["dec.%d.slf.attn" % l] = (B, H, L, L)
        shapes["dec.%d.enc.attn" % l] = (B, H, L, T)
        shapes["dec.%d.slf.proj" % l] = shapes["dec.%d.enc.proj" % l] = shapes["dec.%d.ffn" % l] = (B, L, 128)
    sites = model.dropout_sites
    assert sorted(sites) == sorted(shapes)
    step = model.dropout_state.step_tensor(src.device)
    masks = {}
    for name, shp in shapes.items():
        d = ops.Drop(0.35, sites[name], model.dropout_state.seed, step)
        if name.endswith(".attn"):
            masks[name] = ops.attn_keep_mask(*shp, d, DEV).cpu()
        else:
            masks[name] = ops.dropout_keep_mask(int(np.prod(shp)), d, DEV).cpu().view(shp)
    plan = am.DropoutPlan("injected", masks)
    logits_ref, loss_ref, _, _, grads_ref = otrain.loss_and_grads(sd, cfg, batch[1:], False, plan)
    assert sorted(set(plan.visited)) == sorted(shapes)
    assert rel_err(pred, logits_ref) <= 1e-4
    assert abs(float(loss) - float(loss_ref)) <= 2e-5 * abs(float(loss_ref))
    for k, p in model.named_parameters():
        if k in grads_ref:
            assert rel_err(p.grad, grads_ref[k]) <= 1e-3, k


def test_graphed_step_equals_eager_step():
    """The CUDA-graph replay of the whole step must give the same losses and weights as the eager loop."""
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200.utils import synthetic
    cfg, lda, sd, _ = timit_setup(dropout=0.35)
    batches = synthetic.batches(3, 4, seed=7, pad_to="set")
    results = []
    for graphed in (False, True):
        model = build(model_kwargs(cfg), sd, lda, seed=3)
        opt = pk.ScheduledOptim(pk.FusedAdam(model.parameters(), betas=(0.9, 0.999), eps=1e-8), 1e-3, 25000)
        gs = pk.GraphedTrainStep(model, opt, batches[0]) if graphed else None

        class Loader(list):
            mode = "drop"
        losses = [pk.train_epoch(model, Loader([b]), None, mode="train", optimizer=opt, graphed=gs)[0] for b in batches]
        results.append((losses, {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}))
    (l0, s0), (l1, s1) = results
    assert np.allclose(l0, l1, rtol=1e-6), (l0, l1)
    for k in s0:
        assert torch.allclose(s0[k], s1[k], rtol=1e-5, atol=1e-7), k


def test_no_cpu_fallback():
    import pytorch_kaldi_asr_b200 as pk
    cfg, lda, sd, batch = timit_setup(B=2)
    import pytorch_kaldi_asr_b200.transformer.Models as M
    model = pk.Transformer(lda_mat=lda, **model_kwargs(cfg))
    with pytest.raises(RuntimeError):
        model(torch.from_numpy(batch[1]), torch.from_numpy(batch[2]), torch.from_numpy(batch[3][:, :-1]),
              torch.from_numpy(batch[4][:, :-1]))
