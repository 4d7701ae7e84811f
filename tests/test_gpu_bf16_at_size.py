"""bf16 tensor-core path against the fp32 oracle with the ReLU gate pattern of the GPU run injected into the oracle, and
BASELINE config 5 at its real size (`-m gpu`).

Why gate injection (SURVEY.md 8d asks for gradient cosine >= 0.999 per tensor): a pre-activation that is zero to bf16
precision can land on the other side of a ReLU than in the fp32 oracle.  Its forward contribution is negligible (the
value is ~0 either way) but the unit's whole gradient is switched, so a plain comparison measures the number of flipped
gates, not the arithmetic.  Here the oracle computes `x * gate_gpu` at every ReLU site (oracle.DropoutPlan.relu): both
sides then differentiate the same piecewise-linear function and the remaining difference is bf16 rounding alone --
held to cosine >= 0.999 for every parameter tensor, the bound SURVEY states.  The un-injected comparison stays in
tests/test_gpu_tc.py with its looser, documented bound.

Config 5 (BASELINE configs[4]; reference class T/Models.py:67-124 `Encoder` + T/Models.py:169-231 `Decoder`): d_model 512,
12 + 6 layers, H = 8, d_k = d_v = 64, T ~ 1500, band (-100, 0) and the band opened wide.  At this size the GEMMs take the
128-column-chunk / grid.y plans and attention runs 12-13 key tiles per query tile with band-tile skipping, none of which
the small-size tests reach.  Seeded constructors on both sides, no stored weights.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import acoustic_model as am          # noqa: E402
from oracle import train_step as otrain          # noqa: E402

DEV = "cuda"


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-6))


def cosine(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def relu_sites(cfg):
    """ReLU call order of the tensor-core path == site names of the oracle."""
    if cfg.get("encoder_type", "tdnn") == "tdnn":
        enc = ["enc.tdnn.%d" % i for i in range(len(cfg["tdnn_contexts"]))]
    else:
        enc = ["enc.%d.ffn" % l for l in range(cfg["en_layers"])]
    return enc + ["dec.%d.ffn" % l for l in range(cfg["de_layers"])]


def gpu_bf16_forward_backward(cfg, sd, batch, lda=None):
    """-> (logits, loss, {name: grad}, {relu site: gate u8}) of one bf16 forward/backward in eval mode (no dropout)."""
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200 import ops
    kw = {k: v for k, v in cfg.items() if k != "encoder_type"} if cfg.get("encoder_type", "tdnn") == "tdnn" else dict(cfg)
    model = pk.Transformer(lda_mat=lda, **kw)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    pk.set_compute_mode("bf16")
    ops.GATE_TAP = tap = []
    try:
        src, smask, tgt, tmask = pk.train._to_device(batch, DEV)
        pred = model(src, smask, tgt[:, :-1], tmask[:, :-1])
        loss, _ = pk.get_performance(None, pred, tgt[:, 1:], smoothing=False)
        loss.backward()
        torch.cuda.synchronize()
    finally:
        ops.GATE_TAP = None
        pk.set_compute_mode("fp32")
    names = relu_sites(cfg)
    assert len(tap) == len(names), "expected %d ReLU layers on the tensor-core path, saw %d" % (len(names), len(tap))
    gates = {n: (y > 0).to(torch.uint8).cpu() for n, y in zip(names, tap)}
    grads = {k: p.grad.detach().float().cpu() for k, p in model.named_parameters() if p.grad is not None}
    return pred.detach().float().cpu(), float(loss.detach()), grads, gates


def compare(cfg, sd, batch, lda, logit_tol, cos_bound, label):
    pred, loss, grads, gates = gpu_bf16_forward_backward(cfg, sd, batch, lda)
    plan = am.DropoutPlan("off", gates=gates)
    logits_ref, loss_ref, _, _, grads_ref = otrain.loss_and_grads(sd, cfg, batch[1:], False, plan)
    assert plan.relu_sites == relu_sites(cfg)
    err = rel_err(pred, logits_ref)
    lerr = abs(loss - float(loss_ref)) / abs(float(loss_ref))
    cos = {k: cosine(grads[k], grads_ref[k]) for k in grads_ref}
    worst = min(cos, key=cos.get)
    print("[%s] logits rel err %.3e, loss rel err %.3e, worst gradient cosine %.5f (%s), %d tensors"
          % (label, err, lerr, cos[worst], worst, len(cos)))
    assert set(grads) == set(grads_ref)
    assert err <= logit_tol, "logits: %g" % err
    assert lerr <= 1e-2, "loss: %g" % lerr
    bad = {k: round(v, 5) for k, v in cos.items() if not v >= cos_bound}
    assert not bad, "gradient cosines below %.4f with injected gates: %r" % (cos_bound, bad)


def test_timit_model_bf16_gate_injected_gradient_cosine_0999():
    """TIMIT example model (TDNN encoder), B = 6: every parameter gradient within cosine 0.999 of the oracle once the six
    TDNN ReLU layers and the three FFN ReLUs use the GPU's gate pattern (SURVEY 8d bound, not loosened)."""
    from pytorch_kaldi_asr_b200.utils import synthetic
    cfg = am.example_config(en_dropout=0.0, de_dropout=0.0)
    lda = synthetic.lda_matrix()
    sd = am.init_state_dict(cfg, lda, seed=0)
    batch = synthetic.batches(1, 6, seed=1234)[0]
    compare(cfg, sd, batch, lda, logit_tol=2e-2, cos_bound=0.999, label="timit")


def test_attention_encoder_small_bf16_gate_injected_gradient_cosine_0999():
    from pytorch_kaldi_asr_b200.utils import synthetic
    cfg = am.example_config(en_dropout=0.0, de_dropout=0.0, encoder_type="attention", en_layers=2, de_layers=2,
                            en_d_model=128, de_d_model=128, n_head=2, encoder_sub_sequence=(-30, 0))
    sd = am.init_state_dict(cfg, None, seed=0)
    batch = synthetic.batches(1, 4, seed=99, min_len=140, max_len=300, mean_len=220, std_len=50)[0]
    compare(cfg, sd, batch, None, logit_tol=3e-2, cos_bound=0.999, label="attention-small")


def cfg5_config(band):
    return am.example_config(en_dropout=0.0, de_dropout=0.0, encoder_type="attention", en_layers=12, de_layers=6,
                             en_d_model=512, de_d_model=512, n_head=8, d_k=64, d_v=64, encoder_max_len=1600,
                             decoder_max_len=200, encoder_sub_sequence=band, decoder_sub_sequence=(-20, 0))


@pytest.mark.parametrize("band,B", [((-100, 0), 2), ((-1600, 1600), 1)])
def test_cfg5_at_real_size_bf16_vs_oracle(band, B):
    """BASELINE config 5 at size: 12-layer self-attention encoder + 6-layer decoder, d_model 512, H = 8, T ~ 1500.
    Stated bf16 tolerances (the same as at the small sizes): logits within 2e-2 of the logit scale, summed
    loss within 1e-2 relative, every parameter gradient within cosine 0.999 of the fp32 oracle (gate-injected).  Measured
    on B200 (round 2): logits 7.9e-3 / 9.0e-3, loss 5e-5 / 4e-4, worst cosine 0.99991 over the 280 tensors.  The wide band
    runs one utterance (the oracle materialises [H, T, T] probabilities)."""
    from pytorch_kaldi_asr_b200.utils import synthetic
    cfg = cfg5_config(band)
    sd = am.init_state_dict(cfg, None, seed=0)
    batch = synthetic.batches(1, B, seed=555, mean_len=1500.0, std_len=100.0, min_len=1200, max_len=1599, label_div=10,
                              max_labels=198)[0]
    assert batch[1].shape[1] >= 1200
    compare(cfg, sd, batch, None, logit_tol=2e-2, cos_bound=0.999, label="cfg5 band %r B=%d T=%d" % (band, B, batch[1].shape[1]))
