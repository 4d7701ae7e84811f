"""`-m gpu`: the BatchLoader -> pinned staging -> `train_epoch` feed must be transparent: feeding
the loader (its `next_into` path writes straight into the prefetcher's pinned buffers, which are recycled while earlier
H2D copies are in flight) gives exactly the losses and weights of feeding the same batches as plain numpy tuples.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import acoustic_model as am                      # noqa: E402


class ListLoader(list):
    mode = "drop"


def _run(feed_epochs, cfg, lda, sd, pad_to):
    import pytorch_kaldi_asr_b200 as pk
    model = pk.Transformer(lda_mat=lda, seed=3, **{k: v for k, v in cfg.items() if k != "encoder_type"})
    model.load_state_dict(sd, strict=True)
    model = model.to("cuda")
    opt = pk.ScheduledOptim(pk.FusedAdam(model.parameters(), betas=(0.9, 0.999), eps=1e-8), 1e-3, 25000)
    out = [pk.train_epoch(model, feed, None, mode="train", optimizer=opt) for feed in feed_epochs]
    return out, {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, model


@pytest.mark.parametrize("pad_to", ["dataset", "batch"])
def test_loader_feed_equals_tuple_feed(tmp_path, pad_to):
    from pytorch_kaldi_asr_b200.utils import kaldi_ark, synthetic
    from pytorch_kaldi_asr_b200.utils.BatchLoader import BatchLoader
    cfg = am.example_config(en_dropout=0.35, de_dropout=0.35)
    lda = synthetic.lda_matrix(40, 1, 0)
    sd = am.init_state_dict(cfg, lda, seed=0)
    feats, labels = synthetic.utterances(42, np.random.RandomState(11), mean_len=200.0, max_len=300)
    keys = ["u%02d" % i for i in range(len(feats))]
    kaldi_ark.write_ark_scp(str(tmp_path / "f.ark"), str(tmp_path / "f.scp"), zip(keys, feats))
    scp = kaldi_ark.read_scp(str(tmp_path / "f.scp"))
    triples = [(k, scp[k], lab) for k, lab in zip(keys, labels)]

    def make():
        return BatchLoader(triples, 4, print_info=False, seed=5, pad_to=pad_to, pad_multiple=8)
    recorder = make()
    epochs = [ListLoader(recorder), ListLoader(recorder)]                 # the batches of epoch 1 and epoch 2
    assert len(epochs[0]) == 10 and [b[0] for b in epochs[0]] != [b[0] for b in epochs[1]]
    by_tuples, w_tuples, _ = _run(epochs, cfg, lda, sd, pad_to)
    live = make()
    by_loader, w_loader, model = _run([live, live], cfg, lda, sd, pad_to)
    # same kernels on the same bytes in the same order; a recycled-too-early staging buffer would show up as a gross
    # difference, the tolerance is the one of test_graphed_step_equals_eager_step
    assert np.allclose(np.array(by_loader), np.array(by_tuples), rtol=1e-6), (by_loader, by_tuples)
    for k in w_tuples:
        assert torch.allclose(w_tuples[k], w_loader[k], rtol=1e-5, atol=1e-7), k

    # evaluation: mode 'all' adds the 2-utterance tail batch (a different batch shape through the same staging slots)
    import pytorch_kaldi_asr_b200 as pk
    rec = make()
    rec.mode = "all"
    tuples = ListLoader(rec)
    assert len(tuples) == 11 and len(tuples[-1][0]) == 2
    ev_tuples = pk.train_epoch(model, tuples, None, mode="eval", batch_eval=100)
    ev_loader = pk.train_epoch(model, make(), None, mode="eval", batch_eval=100)
    assert np.allclose(np.array(ev_loader), np.array(ev_tuples), rtol=1e-6), (ev_loader, ev_tuples)
