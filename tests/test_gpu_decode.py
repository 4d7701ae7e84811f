"""Beam-search parity on the B200 (`-m gpu`): the KV-cached, device-lattice decoder must emit the SAME token sequences
as the reference's no-cache host loop (golden fixtures from the real reference + the CPU oracle), fp32 path.

Token equality is demanded whenever the oracle's smallest top-k gap along the search exceeds 1e-4 (far above the fp32
summation-order noise of ~1e-6); scores must agree to 1e-4 absolute.  Any flip is reported with its margin.
"""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import golden_state_dict, load_golden          # noqa: E402
from oracle import acoustic_model as am                      # noqa: E402
from oracle import beam_decode as obd                        # noqa: E402

DEV = "cuda"
SMALL = dict(n_src_dim=8, n_tgt_vocab=11, encoder_max_len=40, decoder_max_len=24, src_fold=1,
             encoder_sub_sequence=(-100, 0), decoder_sub_sequence=(-3, 0), en_layers=2, de_layers=2, n_head=2,
             en_d_model=32, de_d_model=32, d_k=16, d_v=16, en_dropout=0.0, de_dropout=0.0,
             tdnn_contexts=[[-1, 0, 1], [-3, 0, 3]])


def opt(beam, max_len, nbest, **kw):
    return types.SimpleNamespace(use_gpu=True, beam_size=beam, max_token_seq_len=max_len, nbest=nbest, **kw)


def test_lattice_known_answer_vector_through_the_kernel():
    """The reference's only shipped KAT (T/Lattice.py:109-130) driven through pka_beam_advance."""
    from pytorch_kaldi_asr_b200.transformer.Lattice import Lattice
    lat = Lattice(10, 3)
    lat.advance(np.array([[-99, -99, -99, -4, -3, -2, -1]] * 3))
    lat.advance(np.array([[-99, -99, -99, -1.5, -2, -3, -4], [-99, -99, -99, -1.5, -3, -4, -2],
                          [-99, -99, -99, -1.5, -4, -3, -2]]))
    lat.advance(np.array([[-99, -99, -99, -1.5, -2, -3, -4]]))
    assert lat.done
    results, weights = lat.get_results()
    assert results == [[2, 6, 3], [2, 5, 3], [2, 6, 4, 3]]
    assert weights == [-2.5, -3.5, -4.5]
    assert lat.edges == [[-1, 2, 0], [0, 6, -1.0], [0, 5, -2.0], [0, 4, -3.0], [1, 3, -2.5], [1, 4, -3.0], [2, 3, -3.5],
                         [5, 3, -4.5]]
    g = load_golden("semantics")
    assert np.array_equal(np.asarray(lat.edges, dtype=np.float64), g["lattice_edges"])


@pytest.mark.parametrize("beam,nbest,max_len", [(4, 2, 12), (1, 1, 9)])
@pytest.mark.parametrize("use_graph", [False, True])
def test_translate_batch_vs_reference_golden(beam, nbest, max_len, use_graph):
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200.decode import translate_batch
    g = load_golden("decode_small")
    model = pk.Transformer(lda_mat=g["lda_mat"], **SMALL)
    model.load_state_dict(golden_state_dict(g))
    model = model.to(DEV)
    batch = (None, g["src"], g["src_mask"], g["tgt"], None)
    hyps, weights = translate_batch(model, batch, opt(beam, max_len, nbest, use_graph=use_graph), None)
    tag = "beam%d." % beam
    for u in range(len(hyps)):
        want = [g[tag + "hyp.%d.%d" % (u, j)].tolist() for j in range(int(g[tag + "n_hyp.%d" % u]))]
        assert hyps[u] == want, (u, hyps[u], want)
        np.testing.assert_allclose(np.asarray(weights[u]), g[tag + "weights.%d" % u], rtol=0, atol=1e-4)


@pytest.mark.parametrize("force", [False, True])
def test_translate_batch_timit_config_vs_oracle(force):
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200.decode import translate_batch
    from pytorch_kaldi_asr_b200.utils import synthetic
    cfg = am.example_config(en_dropout=0.0, de_dropout=0.0)
    lda = synthetic.lda_matrix()
    sd = am.init_state_dict(cfg, lda, seed=0)
    sd["decoder.tgt_word_proj.linear.weight"] = sd["decoder.tgt_word_proj.linear.weight"] * 4.0   # clearer margins
    batch = synthetic.batches(1, 5, seed=99, mean_len=150, std_len=40, min_len=90, max_len=220)[0]
    beam, max_len, nbest = 10, 14, 3
    ref_h, ref_w, lats, n_steps = obd.translate_batch(sd, cfg, batch[1], batch[2], beam, max_len, nbest,
                                                      force_full_length=force, return_lattices=True)
    model = pk.Transformer(lda_mat=lda, **{k: v for k, v in cfg.items() if k != "encoder_type"})
    model.load_state_dict(sd)
    model = model.to(DEV)
    hyps, weights = translate_batch(model, batch, opt(beam, max_len, nbest, force_full_length=force), None)
    flips = []
    for u in range(len(hyps)):
        np.testing.assert_allclose(np.asarray(weights[u]), np.asarray(ref_w[u]), rtol=0, atol=1e-4)
        if hyps[u] != ref_h[u]:
            flips.append((u, lats[u].min_gap))
    # a flip is only acceptable where the oracle itself had a near-tie
    assert all(gap <= 1e-4 for _, gap in flips), "token flips with margins %r" % flips
    if force:
        assert n_steps == max_len and all(len(h[0]) == max_len + 1 for h in hyps)


def _golden_run(name, dec_band, beam, nbest, max_len, **extra):
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200.decode import translate_batch
    g = load_golden(name)
    model = pk.Transformer(lda_mat=g["lda_mat"], **dict(SMALL, decoder_sub_sequence=dec_band))
    model.load_state_dict(golden_state_dict(g))
    model = model.to(DEV)
    batch = (None, g["src"], g["src_mask"], g["tgt"], None)
    hyps, weights = translate_batch(model, batch, opt(beam, max_len, nbest, **extra), None)
    tag = "beam%d." % beam
    for u in range(len(hyps)):
        want = [g[tag + "hyp.%d.%d" % (u, j)].tolist() for j in range(int(g[tag + "n_hyp.%d" % u]))]
        assert hyps[u] == want, (u, hyps[u], want)
        np.testing.assert_allclose(np.asarray(weights[u]), g[tag + "weights.%d" % u], rtol=0, atol=1e-4)


@pytest.mark.parametrize("beam,nbest,max_len", [(4, 2, 12), (1, 1, 9)])
def test_noncausal_decoder_band_vs_reference_golden(beam, nbest, max_len):
    """Decoder band (-3, 1): not cacheable, decoded by the no-cache path (full-prefix decoder re-runs, like the
    reference L/decode.py:81-87) -- tokens and scores of the real reference (tests/golden/make_golden.py)."""
    _golden_run("decode_small_band1", (-3, 1), beam, nbest, max_len)


@pytest.mark.parametrize("beam,nbest,max_len", [(4, 2, 12)])
def test_no_cache_path_on_a_causal_band_matches_the_reference_too(beam, nbest, max_len):
    """`opt.use_cache = False` forces the no-cache path on the causal band the KV-cached path is tested on above."""
    _golden_run("decode_small", (-3, 0), beam, nbest, max_len, use_cache=False)


def test_no_cache_path_chunks_utterances():
    """The no-cache path bounds the replicated encoder memory by decoding utterance chunks; chunked == unchunked."""
    import pytorch_kaldi_asr_b200 as pk
    from pytorch_kaldi_asr_b200 import decode as D
    g = load_golden("decode_small_band1")
    model = pk.Transformer(lda_mat=g["lda_mat"], **dict(SMALL, decoder_sub_sequence=(-3, 1)))
    model.load_state_dict(golden_state_dict(g))
    model = model.to(DEV)
    batch = (None, g["src"], g["src_mask"], g["tgt"], None)
    whole = D.translate_batch(model, batch, opt(4, 8, 2), None)
    bd = D.BeamDecoder(model, 2, g["src"].shape[1], 4, 8, use_cache=False)
    with torch.no_grad():
        enc, fmask = model.encode(torch.from_numpy(g["src"]).to(DEV), torch.from_numpy(g["src_mask"]).to(DEV))
    bd.reset(enc[1:3], fmask[1:3])
    bd.run()
    part = bd.results(2)
    assert part[0] == whole[0][1:3]
    for a, b in zip(part[1], whole[1][1:3]):
        np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=0, atol=1e-5)
