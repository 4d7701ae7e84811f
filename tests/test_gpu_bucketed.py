"""Bucketed padding (`-m gpu`): a batch padded to bucket_length(longest) -- at least the TDNN stack's 16 frames of right
context behind every utterance (L/pytorch/TDNN.py:25-28, SURVEY 8e caveat) -- gives the results of the reference's
whole-set padding on every real position: logits bit for bit, loss and gradients to fp32 summation-order noise."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import acoustic_model as am          # noqa: E402

DEV = "cuda"


def repad(batch, T, L1):
    def grow(x, n):
        out = np.zeros((x.shape[0], n) + x.shape[2:], dtype=x.dtype)
        out[:, :x.shape[1]] = x
        return out
    return (batch[0], grow(batch[1], T), grow(batch[2], T), grow(batch[3], L1), grow(batch[4], L1))


def run(model, batch, pk):
    from pytorch_kaldi_asr_b200 import ops
    for p in model.parameters():
        p.grad = None
    src, smask, tgt, tmask = pk.train._to_device(batch, DEV)
    tgt_in, goal, tmask_in = ops.split_targets(tgt, tmask)
    pred = model(src, smask, tgt_in, tmask_in)
    loss, _ = ops.cross_entropy_sum(pred.view(-1, pred.size(-1)), goal.view(-1), False)
    loss.backward()
    torch.cuda.synchronize()
    return pred.detach(), float(loss.detach()), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_bucket_padding_equals_whole_set_padding_on_real_positions(mode):
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200.utils import synthetic
    cfg = am.example_config(en_dropout=0.0, de_dropout=0.0)
    lda = synthetic.lda_matrix()
    sd = am.init_state_dict(cfg, lda, seed=0)
    pool = synthetic.batches(8, 8, seed=77, pad_to="bucket")
    set_T = max(int(b[2].sum(axis=1).max()) for b in pool)
    batch = min(pool, key=lambda b: b[1].shape[1])                  # the shortest bucket
    T_b, L_b = batch[1].shape[1], batch[3].shape[1]
    lens = batch[2].sum(axis=1)
    assert T_b == synthetic.bucket_length(int(lens.max()), set_T) and T_b < set_T and T_b - int(lens.max()) >= 16
    whole = repad(batch, set_T, 100)                                # the reference's whole-set shape
    model = pk.Transformer(lda_mat=lda, **{k: v for k, v in cfg.items() if k != "encoder_type"})
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    pk.set_compute_mode(mode)
    try:
        p_b, l_b, g_b = run(model, batch, pk)
        p_w, l_w, g_w = run(model, whole, pk)
    finally:
        pk.set_compute_mode("fp32")
    tok = torch.as_tensor(batch[4][:, :-1]).bool().to(DEV)          # real decoder positions
    a, b = p_b[tok], p_w[:, :L_b - 1][tok]
    assert torch.equal(a, b), "logits on real positions differ: max %g" % float((a - b).abs().max())
    assert abs(l_b - l_w) <= 1e-6 * abs(l_w)
    for k in g_b:
        rel = float((g_b[k] - g_w[k]).abs().max() / g_w[k].abs().max().clamp_min(1e-20))
        assert rel <= 2e-5, "%s: %g" % (k, rel)


def test_graphed_step_handles_several_bucket_shapes_like_the_eager_step():
    """Same weights after stepping through batches of different bucket shapes with per-shape CUDA graphs and eagerly."""
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200.utils import synthetic
    cfg = am.example_config(en_dropout=0.35, de_dropout=0.35)
    lda = synthetic.lda_matrix()
    sd = am.init_state_dict(cfg, lda, seed=0)
    pool = synthetic.batches(4, 6, seed=5, pad_to="bucket")
    assert len({b[1].shape[1] for b in pool}) >= 2
    results = []
    pk.set_compute_mode("bf16")
    try:
        for graphed in (False, True):
            model = pk.Transformer(lda_mat=lda, seed=11, **{k: v for k, v in cfg.items() if k != "encoder_type"})
            model.load_state_dict(sd)
            model = model.to(DEV).train()
            opt = pk.ScheduledOptim(pk.FusedAdam(model.parameters()), 1e-3, 25000)
            gs = pk.GraphedTrainStep(model, opt) if graphed else None

            class Loader(list):
                mode = "drop"
            loss, acc = pk.train_epoch(model, Loader(pool + pool[::-1]), None, mode="train", optimizer=opt, graphed=gs)
            torch.cuda.synchronize()
            results.append((loss, acc, opt.optimizer.flat_param.clone(), int(opt.optimizer.dev_state[0]), opt.n_current_steps))
    finally:
        pk.set_compute_mode("fp32")
    (l0, a0, w0, t0, n0), (l1, a1, w1, t1, n1) = results
    assert t0 == t1 == 8 and n0 == n1 == 8
    assert abs(l0 - l1) <= 1e-5 * abs(l0) and a0 == a1
    assert float((w0 - w1).abs().max()) <= 1e-5
