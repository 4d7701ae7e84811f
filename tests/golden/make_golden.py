#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the REAL reference (boji123/pytorch-kaldi-asr) in place.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference is imported unmodified from /root/reference with the two shims of SURVEY.md Appendix B (a stub
`kaldi_io` module, and Bottle.forward `view`->`reshape`, needed on torch>=2 and numerically identical).  Nothing from
the reference is copied; only its *outputs* on seeded synthetic inputs are stored (float32/int64/float64 arrays in
compressed .npz files), together with the weights its own constructors produced under torch.manual_seed.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
R = os.environ.get("PKA_REFERENCE", "/root/reference")


def import_reference():
    sys.path[:0] = [R + "/pytorch",
                    R + "/project/attention-transformer-timit/local/pytorch",
                    R + "/project/attention-transformer-timit/local"]
    sys.modules.setdefault("kaldi_io", types.ModuleType("kaldi_io"))
    import transformer.Modules as M

    def bottle_forward(self, input):
        if len(input.size()) <= 2:
            return super(M.Bottle, self).forward(input)
        size = input.size()[:2]
        out = super(M.Bottle, self).forward(input.reshape(size[0] * size[1], -1))
        return out.view(size[0], size[1], -1)

    M.Bottle.forward = bottle_forward
    import train, decode                                     # noqa: E401  (reference L/train.py, L/decode.py)
    import transformer.Models as Models
    import transformer.Lattice as Lattice
    import transformer.Optim as Optim
    import TDNN
    from utils import instances_handler
    return types.SimpleNamespace(train=train, decode=decode, Models=Models, Lattice=Lattice, Optim=Optim, TDNN=TDNN,
                                 M=M, ih=instances_handler)


def synth_batch(rng, n_utt, t_lo, t_hi, feat, vocab, ref):
    feats, labels = [], []
    for _ in range(n_utt):
        t = int(rng.randint(t_lo, t_hi + 1))
        feats.append(rng.randn(t, feat).astype(np.float32))
        n_lab = max(2, t // 4)
        labels.append(np.concatenate([[2], rng.randint(4, vocab, size=n_lab), [3]]).astype(np.int64))
    src, src_mask = ref.ih.pad_to_longest(feats)
    tgt, tgt_mask = ref.ih.pad_to_longest(labels)
    keys = ["utt%d" % i for i in range(n_utt)]
    return keys, src.astype(np.float32), src_mask, tgt.astype(np.int64), tgt_mask


def sd_numpy(model, prefix="sd."):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


SMALL = dict(n_src_dim=8, n_tgt_vocab=11, encoder_max_len=40, decoder_max_len=24, src_fold=1,
             encoder_sub_sequence=(-100, 0), decoder_sub_sequence=(-3, 0), en_layers=2, de_layers=2, n_head=2,
             en_d_model=32, de_d_model=32, d_k=16, d_v=16, en_dropout=0.0, de_dropout=0.0,
             tdnn_contexts=[[-1, 0, 1], [-3, 0, 3]])


def build(ref, cfg, seed):
    torch.manual_seed(seed)
    s = cfg["n_src_dim"] * cfg["src_fold"] * 5
    lda = (np.random.RandomState(seed).randn(s, s + 1) * 0.1).astype(np.float32)
    return ref.Models.Transformer(cfg["n_src_dim"], cfg["n_tgt_vocab"], lda_mat=lda,
                                  **{k: v for k, v in cfg.items() if k not in ("n_src_dim", "n_tgt_vocab")}), lda


def model_fwd_bwd(ref, cfg, seed, name, n_utt=3, t_lo=9, t_hi=21):
    model, lda = build(ref, cfg, seed)
    model.eval()
    rng = np.random.RandomState(100 + seed)
    keys, src, src_mask, tgt, tgt_mask = synth_batch(rng, n_utt, t_lo, t_hi, cfg["n_src_dim"], cfg["n_tgt_vocab"], ref)
    out = dict(sd_numpy(model), lda_mat=lda, src=src, src_mask=src_mask, tgt=tgt, tgt_mask=tgt_mask)
    tsrc, tsm = torch.from_numpy(src), torch.from_numpy(src_mask)
    ttgt, ttm = torch.from_numpy(tgt), torch.from_numpy(tgt_mask)
    goal, tgt_in, tgt_in_mask = ttgt[:, 1:], ttgt[:, :-1], ttm[:, :-1]
    for smoothing in (False, True):
        model.zero_grad()
        pred = model(tsrc, tsm, tgt_in, tgt_in_mask)
        loss, n_correct = ref.train.get_performance(None, pred, goal, smoothing=smoothing)
        loss.backward()
        tag = "smooth." if smoothing else "plain."
        out[tag + "loss"] = np.float32(loss.item())
        out[tag + "n_correct"] = np.int64(int(n_correct))
        for k, p in model.named_parameters():
            if p.grad is not None:
                out[tag + "grad." + k] = p.grad.detach().numpy().copy()
    out["logits"] = pred.detach().numpy()
    out["n_words"] = np.int64(int(goal.ne(0).sum()))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return model, (keys, src, src_mask, tgt, tgt_mask)


def attention_encoder_golden(ref, name):
    """The unwired self-attention Encoder (T/Models.py:67-124), driven stand-alone, plus a Decoder on top."""
    torch.manual_seed(7)
    enc = ref.Models.Encoder(n_src_dim=8, encoder_max_len=40, n_layers=2, n_head=2, sub_sequence=(-4, 1),
                             d_k=16, d_v=16, d_model=32, d_inner_hid=32, dropout=0.0)
    dec = ref.Models.Decoder(n_tgt_vocab=11, decoder_max_len=24, n_layers=1, n_head=2, sub_sequence=(-3, 0),
                             d_k=16, d_v=16, en_d_model=32, de_d_model=32, d_inner_hid=32, dropout=0.0)
    enc.eval(); dec.eval()
    rng = np.random.RandomState(107)
    keys, src, src_mask, tgt, tgt_mask = synth_batch(rng, 3, 9, 21, 8, 11, ref)
    tsrc, tsm = torch.from_numpy(src), torch.from_numpy(src_mask)
    ttgt, ttm = torch.from_numpy(tgt), torch.from_numpy(tgt_mask)
    enc_out, = enc(tsrc, tsm)
    logits, = dec(ttgt[:, :-1], ttm[:, :-1], tsm, enc_out)
    loss = ref.train.cal_loss(logits.reshape(-1, logits.size(2)), ttgt[:, 1:].contiguous(), False)
    loss.backward()
    out = dict(src=src, src_mask=src_mask, tgt=tgt, tgt_mask=tgt_mask, enc_out=enc_out.detach().numpy(),
               logits=logits.detach().numpy(), loss=np.float32(loss.item()))
    for k, v in enc.state_dict().items():
        out["sd.encoder." + k] = v.numpy()
    for k, v in dec.state_dict().items():
        out["sd.decoder." + k] = v.numpy()
    for k, p in enc.named_parameters():
        if p.grad is not None:
            out["grad.encoder." + k] = p.grad.numpy().copy()
    for k, p in dec.named_parameters():
        if p.grad is not None:
            out["grad.decoder." + k] = p.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


class ListLoader(list):
    mode = "drop"


def train_steps_golden(ref, name):
    cfg = dict(SMALL)
    model, lda = build(ref, cfg, 3)
    rng = np.random.RandomState(303)
    batches = [synth_batch(rng, 4, 9, 21, cfg["n_src_dim"], cfg["n_tgt_vocab"], ref) for _ in range(4)]
    out = dict(sd_numpy(model, "sd0."), lda_mat=lda)
    opt = ref.Optim.ScheduledOptim(torch.optim.Adam(model.parameters(), betas=(0.9, 0.999), eps=1e-8), 2e-3, 10)
    per_step = []
    for i, b in enumerate(batches):
        loss_per_word, acc = ref.train.train_epoch(model, ListLoader([b]), None, mode="train", optimizer=opt)
        per_step.append([loss_per_word, acc])
        for k, v in zip("src src_mask tgt tgt_mask".split(), b[1:]):
            out["batch%d.%s" % (i, k)] = v
    out["per_step"] = np.asarray(per_step, dtype=np.float64)
    out.update(sd_numpy(model, "sd4."))
    ev = ref.train.train_epoch(model, ListLoader(batches), None, mode="eval", batch_eval=3)
    out["eval_after"] = np.asarray(ev, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def decode_golden(ref, name, dec_band=(-3, 0)):
    """`dec_band` with end > 0 is a non-causal decoder band: the reference decodes it like any other (it re-runs the
    whole decoder over the full prefixes, L/decode.py:81-87); here it pins the no-cache path of decode.py."""
    cfg = dict(SMALL, decoder_sub_sequence=dec_band)
    model, lda = build(ref, cfg, 5)
    # sharpen the output layer so hypotheses differ clearly and some reach EOS within the step budget
    # and give the encoder more say / make EOS competitive so finished hypotheses stay in the beam and compete
    with torch.no_grad():
        w = model.decoder.tgt_word_proj.linear.weight
        w.mul_(6.0)
        model.decoder.enc_dec_projection.linear.weight.mul_(3.0)
        w[3] = 0.5 * (w[8] + w[9])
    model.prob_projection = torch.nn.LogSoftmax(dim=1)
    rng = np.random.RandomState(505)
    batch = synth_batch(rng, 5, 9, 30, cfg["n_src_dim"], cfg["n_tgt_vocab"], ref)
    out = dict(sd_numpy(model), lda_mat=lda, src=batch[1], src_mask=batch[2], tgt=batch[3])
    for beam, nbest, max_len in ((4, 2, 12), (1, 1, 9)):
        opt = types.SimpleNamespace(use_gpu=False, beam_size=beam, max_token_seq_len=max_len, nbest=nbest)
        hyps, weights = ref.decode.translate_batch(model, batch, opt, None)
        tag = "beam%d." % beam
        for u, (h, w) in enumerate(zip(hyps, weights)):
            out[tag + "n_hyp.%d" % u] = np.int64(len(h))
            for j, seq in enumerate(h):
                out[tag + "hyp.%d.%d" % (u, j)] = np.asarray(seq, dtype=np.int64)
            out[tag + "weights.%d" % u] = np.asarray(w, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def semantics_golden(ref, name):
    """Appendix A unit vectors."""
    out = {}
    x = torch.arange(1, 5, dtype=torch.float32).view(1, 4, 1)
    out["concat_1234"] = ref.TDNN.ConcatLayer([-2, -1, 0, 1, 2])(x).numpy()
    xr = torch.from_numpy(np.random.RandomState(1).randn(2, 7, 3).astype(np.float32))
    out["concat_in"] = xr.numpy()
    out["concat_m303"] = ref.TDNN.ConcatLayer([-3, 0, 3])(xr).numpy()
    m = torch.from_numpy(np.array([[1, 1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 0, 0, 0]], dtype=np.uint8))
    fs, fm = ref.Models.fold_seq_and_mask(xr, m, 2)
    out["fold2_seq"], out["fold2_mask"] = fs.numpy(), fm.numpy()
    fs, fm = ref.Models.fold_seq_and_mask(xr, m, 3)
    out["fold3_seq"], out["fold3_mask"] = fs.numpy(), fm.numpy()
    km = torch.from_numpy(np.array([[1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 0, 0]], dtype=np.uint8))
    pad = ref.Models.get_attn_padding_mask(km, km)
    sub = ref.Models.get_attn_subsequent_mask(km, -2, 0)
    out["mask_in"] = km.numpy()
    out["mask_m2_0"] = torch.gt(pad + sub, 0).numpy()
    sub = ref.Models.get_attn_subsequent_mask(km, -1, 2)
    out["mask_m1_2"] = torch.gt(pad + sub, 0).numpy()
    out["posenc_9x6"] = ref.Models.position_encoding_init(9, 6).numpy()
    ln = ref.M.LayerNormalization(5)
    with torch.no_grad():
        ln.a_2.copy_(torch.tensor([1.0, 2.0, 0.5, -1.0, 1.5]))
        ln.b_2.copy_(torch.tensor([0.0, 0.1, -0.2, 0.3, 0.0]))
    z = torch.from_numpy(np.random.RandomState(2).randn(2, 3, 5).astype(np.float32))
    out["ln_in"], out["ln_out"] = z.numpy(), ln(z).detach().numpy()
    out["ln_len1_out"] = ln(z[:, :1]).detach().numpy()
    out["ln_a"], out["ln_b"] = ln.a_2.detach().numpy(), ln.b_2.detach().numpy()
    # scaled dot-product attention incl. a fully masked row (T/Modules.py:75-97)
    att = ref.M.ScaledDotProductAttention(d_model=32, attn_dropout=0.0)
    r = np.random.RandomState(3)
    q, k, v = [torch.from_numpy(r.randn(2, 4, 8).astype(np.float32)).requires_grad_(True) for _ in range(3)]
    msk = torch.zeros(2, 4, 4, dtype=torch.bool)
    msk[0, 1, :] = True
    msk[1, :, 2:] = True
    o, p = att(q, k, v, attn_mask=msk)
    wsum = torch.from_numpy(r.randn(2, 4, 8).astype(np.float32))
    (o * wsum).sum().backward()
    out.update(sdpa_q=q.detach().numpy(), sdpa_k=k.detach().numpy(), sdpa_v=v.detach().numpy(), sdpa_mask=msk.numpy(),
               sdpa_out=o.detach().numpy(), sdpa_probs=p.detach().numpy(), sdpa_w=wsum.numpy(),
               sdpa_dq=q.grad.numpy(), sdpa_dk=k.grad.numpy(), sdpa_dv=v.grad.numpy())
    # loss
    lg = torch.from_numpy(r.randn(6, 7).astype(np.float32))
    goal = torch.tensor([4, 0, 6, 3, 0, 5])
    out["ce_logits"], out["ce_goal"] = lg.numpy(), goal.numpy()
    out["ce_plain"] = np.float32(ref.train.cal_loss(lg, goal, False).item())
    out["ce_smooth"] = np.float32(ref.train.cal_loss(lg, goal, True).item())
    # lattice demo (the only known-answer vector the reference ships, T/Lattice.py:109-130)
    lat = ref.Lattice.Lattice(10, 3)
    a = [[-99, -99, -99, -4, -3, -2, -1]] * 3
    lat.advance(np.array(a))
    lat.advance(np.array([[-99, -99, -99, -1.5, -2, -3, -4], [-99, -99, -99, -1.5, -3, -4, -2],
                          [-99, -99, -99, -1.5, -4, -3, -2]]))
    lat.advance(np.array([[-99, -99, -99, -1.5, -2, -3, -4]]))
    res, w = lat.get_results()
    out["lattice_done"] = np.bool_(lat.done)
    out["lattice_weights"] = np.asarray(w, dtype=np.float64)
    out["lattice_edges"] = np.asarray(lat.edges, dtype=np.float64)
    for i, s in enumerate(res):
        out["lattice_result.%d" % i] = np.asarray(s, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def main():
    ref = import_reference()
    torch.set_num_threads(1)                      # deterministic summation order for the stored floats
    model_fwd_bwd(ref, dict(SMALL), 1, "tdnn_small_fwd_bwd")
    model_fwd_bwd(ref, dict(SMALL, src_fold=2, encoder_max_len=20, decoder_sub_sequence=(-2, 0)), 2,
                  "tdnn_small_fold2_fwd_bwd", n_utt=4, t_lo=11, t_hi=27)
    attention_encoder_golden(ref, "attn_encoder_small")
    train_steps_golden(ref, "train_steps_small")
    decode_golden(ref, "decode_small")
    decode_golden(ref, "decode_small_band1", dec_band=(-3, 1))
    semantics_golden(ref, "semantics")
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print("%-36s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()
