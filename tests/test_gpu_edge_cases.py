"""Edge cases of the path on the B200 (`-m gpu`), each against the CPU oracle on the same padded tensors:
maximum lengths (encoder_max_len frames, decoder_max_len tokens) and one past them, a batch of one, utterances shorter
than the TDNN receptive field, an utterance that is padding only, and the bf16 tensor-core path in training mode with
all 25 dropout sites on (the kernels' Philox bits are materialised and injected into the oracle).
Tolerances: fp32 path as in test_gpu_model.py (logits 1e-4, loss 2e-5 rel., gradients 1e-3 of the tensor scale);
bf16 path as in test_gpu_tc.py (logits 2e-2, loss 1e-2 rel.)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import acoustic_model as am                      # noqa: E402
from oracle import train_step as otrain                      # noqa: E402

DEV = "cuda"


def rel_err(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-6))


def make_batch(src_lens, tgt_lens, T=None, L=None, seed=0, F=40, V=53):
    rng = np.random.RandomState(seed)
    B = len(src_lens)
    T = T or max(max(src_lens), 1)
    L = L or max(tgt_lens) + 2
    src = np.zeros((B, T, F), np.float32)
    smask = np.zeros((B, T), np.uint8)
    tgt = np.zeros((B, L), np.int64)
    tmask = np.zeros((B, L), np.uint8)
    for b, (t, l) in enumerate(zip(src_lens, tgt_lens)):
        src[b, :t] = rng.randn(t, F)
        smask[b, :t] = 1
        seq = [2] + list(rng.randint(4, V - 1, size=l)) + [3]
        tgt[b, :len(seq)] = seq
        tmask[b, :len(seq)] = 1
    return (None, src, smask, tgt, tmask)


def setup(**over):
    from pytorch_kaldi_asr_b200.utils import synthetic
    cfg = am.example_config(en_dropout=0.0, de_dropout=0.0, **over)
    lda = synthetic.lda_matrix(40, 1, 0)
    return cfg, lda, am.init_state_dict(cfg, lda, seed=0)


def run_and_compare(cfg, lda, sd, batch, check_grads=True):
    import pytorch_kaldi_asr_b200 as pk
    logits_ref, loss_ref, nc_ref, _, grads_ref = otrain.loss_and_grads(sd, cfg, batch[1:], smoothing=False)
    model = pk.Transformer(lda_mat=lda, **{k: v for k, v in cfg.items() if k != "encoder_type"})
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    src, smask, tgt, tmask = pk.train._to_device(batch, DEV)
    pred = model(src, smask, tgt[:, :-1], tmask[:, :-1])
    assert torch.isfinite(pred).all()
    assert rel_err(pred, logits_ref) <= 1e-4
    loss, n_correct = pk.get_performance(None, pred, tgt[:, 1:], smoothing=False)
    assert abs(float(loss) - float(loss_ref)) <= 2e-5 * max(abs(float(loss_ref)), 1e-6)
    assert int(n_correct) == nc_ref
    if check_grads:
        loss.backward()
        for k, p in model.named_parameters():
            if k in grads_ref:
                assert torch.isfinite(p.grad).all(), k
                assert rel_err(p.grad, grads_ref[k]) <= 1e-3, k
    return model


def test_maximum_lengths_and_one_past_them():
    """encoder_max_len = 500 frames and decoder_max_len = 100 decoder positions are the position tables' sizes
    (P/run.sh:30,78); both exactly full must work, one more must raise (the reference indexes out of range there)."""
    import pytorch_kaldi_asr_b200 as pk
    cfg, lda, sd = setup()
    batch = make_batch([500, 317], [99, 40], T=500, L=101, seed=1)          # 100 decoder inputs after dropping the last token
    model = run_and_compare(cfg, lda, sd, batch, check_grads=False)
    too_long = make_batch([501], [10], T=501, seed=2)
    src, smask, tgt, tmask = pk.train._to_device(too_long, DEV)
    with pytest.raises((RuntimeError, AssertionError)):
        model(src, smask, tgt[:, :-1], tmask[:, :-1])
    too_many = make_batch([50], [100], T=50, L=102, seed=3)                 # 101 decoder positions
    src, smask, tgt, tmask = pk.train._to_device(too_many, DEV)
    with pytest.raises((RuntimeError, AssertionError)):
        model(src, smask, tgt[:, :-1], tmask[:, :-1])


def test_batch_of_one_and_utterances_shorter_than_the_receptive_field():
    """The TDNN stack sees +-16 frames; utterances of 1, 3 and 7 frames (padded to 9) still match the oracle, and so does
    a batch holding a single utterance."""
    cfg, lda, sd = setup()
    run_and_compare(cfg, lda, sd, make_batch([1, 3, 7, 9], [5, 5, 6, 7], T=9, seed=4))
    run_and_compare(cfg, lda, sd, make_batch([123], [17], seed=5))


def test_utterance_that_is_padding_only():
    """All-zero features with an all-zero pad mask next to a normal utterance: cross-attention rows have no allowed key
    (probabilities re-filled with 0, T/Modules.py:90), outputs stay finite and equal the oracle's."""
    cfg, lda, sd = setup()
    batch = make_batch([60, 0], [9, 4], T=60, seed=6)
    run_and_compare(cfg, lda, sd, batch)


def test_bf16_training_mode_with_all_dropout_sites_vs_oracle_with_injected_masks():
    """Tensor-core path, model.train(), p = 0.35 at all 25 sites: keep bits from the kernels' Philox definition are
    injected into the fp32 oracle; logits / loss agree within the bf16 tolerances."""
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200 import ops
    from pytorch_kaldi_asr_b200.utils import synthetic
    cfg = am.example_config(en_dropout=0.35, de_dropout=0.35)
    lda = synthetic.lda_matrix(40, 1, 0)
    sd = am.init_state_dict(cfg, lda, seed=0)
    batch = synthetic.batches(1, 4, seed=77)[0]
    model = pk.Transformer(lda_mat=lda, seed=5, **{k: v for k, v in cfg.items() if k != "encoder_type"})
    model.load_state_dict(sd)
    model = model.to(DEV).train()
    pk.set_compute_mode("bf16")
    try:
        src, smask, tgt, tmask = pk.train._to_device(batch, DEV)
        pred = model(src, smask, tgt[:, :-1], tmask[:, :-1])
        loss, _ = pk.get_performance(None, pred, tgt[:, 1:], smoothing=False)
        loss.backward()
    finally:
        pk.set_compute_mode("fp32")
    B, T, L, H = src.shape[0], src.shape[1], tgt.shape[1] - 1, cfg["n_head"]
    shapes = {"enc.src": (B, T, 256), "enc.out": (B, T, 256), "dec.emb": (B, L, 128), "dec.out": (B, L, 128)}
    for i in range(6):
        shapes["enc.tdnn.%d" % i] = (B, T, 256)
    for l in range(3):
        shapes["dec.%d.slf.attn" % l] = (B, H, L, L)
        shapes["dec.%d.enc.attn" % l] = (B, H, L, T)
        shapes["dec.%d.slf.proj" % l] = shapes["dec.%d.enc.proj" % l] = shapes["dec.%d.ffn" % l] = (B, L, 128)
    sites = model.dropout_sites
    step = model.dropout_state.step_tensor(src.device)
    masks = {}
    for name, shp in shapes.items():
        d = ops.Drop(0.35, sites[name], model.dropout_state.seed, step)
        masks[name] = ops.attn_keep_mask(*shp, d, DEV).cpu() if name.endswith(".attn") else \
            ops.dropout_keep_mask(int(np.prod(shp)), d, DEV).cpu().view(shp)
    plan = am.DropoutPlan("injected", masks)
    logits_ref, loss_ref, _, _, grads_ref = otrain.loss_and_grads(sd, cfg, batch[1:], False, plan)
    assert rel_err(pred, logits_ref) <= 2e-2
    assert abs(float(loss) - float(loss_ref)) <= 1e-2 * abs(float(loss_ref))
    a = model.decoder.tgt_word_proj.linear.weight.grad.double().flatten().cpu()
    b = grads_ref["decoder.tgt_word_proj.linear.weight"].double().flatten()
    assert float((a @ b) / (a.norm() * b.norm())) >= 0.999
