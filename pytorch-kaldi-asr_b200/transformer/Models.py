"""Top-level acoustic model with the reference's constructor / forward API and state-dict keys (T/Models.py), running
on hand-written sm_100a kernels.  `Transformer` wires the TDNN encoder (`encoder_test`) exactly like the reference;
`encoder_type="attention"` additionally offers the self-attention `Encoder` class that the reference defines but leaves
unwired (T/Models.py:67-124, 242-246) -- BASELINE config 5 needs it.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .. import rng as _rng
from ..TDNN import ConcatLayer, LDALayer, TDNNLayer
from ..utils import constants
from .Layers import DecoderLayer, EncoderLayer
from .Modules import AttnMask
from .Modules import BottleLinear as Linear

LDA_CONCAT_INDEX = [-2, -1, 0, 1, 2]          # T/Models.py:140
CMVN_MODES = {"none": 0, None: 0, False: 0, "mean": 1, "meanvar": 2}


def position_encoding_init(n_position, d_pos_vec):
    """Sinusoid table, row 0 all zeros, sin on even / cos on odd columns (T/Models.py:16-25), vectorised."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    col = np.arange(d_pos_vec, dtype=np.float64)[None, :]
    angle = pos / np.power(10000.0, 2.0 * np.floor(col / 2.0) / d_pos_vec)
    table = np.where(np.arange(d_pos_vec)[None, :] % 2 == 0, np.sin(angle), np.cos(angle))
    table[0] = 0.0
    return torch.from_numpy(table).type(torch.FloatTensor)


def get_attn_padding_mask(seq_q, seq_k):
    """Key-padding part of the mask (T/Models.py:27-36) as a symbolic AttnMask; `seq_k` is the uint8 pad mask."""
    assert seq_q.dim() == 2 and seq_k.dim() == 2
    return AttnMask(key_pad_mask=seq_k, band=None, q_len=seq_q.size(1))


def get_attn_subsequent_mask(seq, start, end):
    """Band part of the mask: query i sees keys i+start..i+end (T/Models.py:38-49).  No host np.triu, no H2D copy."""
    assert seq.dim() == 2
    return AttnMask(key_pad_mask=None, band=(int(start), int(end)), q_len=seq.size(1))


def fold_seq_and_mask(seq, pad_mask, fold):
    """Stack `fold` consecutive frames; a stacked frame is real iff its last constituent is (T/Models.py:51-65).
    Pure view/stride work; `Transformer.forward` instead folds inside the fused front-end kernel."""
    if fold == 1:
        return seq, pad_mask
    if fold < 1:
        raise ValueError("[ERROR] invalid data fold parameter")
    keep = seq.size(1) - seq.size(1) % fold
    seq = seq[:, :keep].contiguous().view(seq.size(0), -1, seq.size(2) * fold)
    return seq, pad_mask[:, fold - 1::fold].contiguous()


def _frozen_table(n, d):
    emb = nn.Embedding(n, d, padding_idx=constants.PAD)
    emb.weight.data = position_encoding_init(n, d)
    emb.weight.requires_grad = False
    return emb


class Encoder(nn.Module):
    """Self-attention encoder (T/Models.py:67-124): Linear(F->D) + pos -> dropout -> N x EncoderLayer -> + pos -> dropout.
    Returns a 1-tuple like the reference."""

    def __init__(self, n_src_dim, encoder_max_len, n_layers=2, n_head=3, sub_sequence=(-1, 1), d_k=64, d_v=64,
                 d_model=256, d_inner_hid=256, dropout=0.1, rng=None):
        super().__init__()
        self.sub = sub_sequence
        self.d_model = d_model
        self.p = float(dropout)
        self._rng = rng or _rng.GLOBAL
        self._site_in = self._rng.site("enc.in")
        self.position_enc = _frozen_table(encoder_max_len, d_model)
        self.trans_pos_enc = _frozen_table(encoder_max_len, d_model)
        self.src_projection = Linear(n_src_dim, d_model, bias=False)
        self.layer_stack = nn.ModuleList([
            EncoderLayer(d_model, d_inner_hid, n_head, d_k, d_v, dropout=dropout, rng=self._rng, site="enc.%d" % i)
            for i in range(n_layers)])
        self._site_out = self._rng.site("enc.out")

    def forward(self, src_seq, src_pad_mask, return_attns=False):
        dev = src_seq.device
        if ops.compute_mode() == "bf16":       # bf16 activation stream: the front-end kernel makes the bf16 copy
            src_seq = ops.frontend(src_seq, None, 1, [0], 0, out_dtype=torch.bfloat16)
        x = self.src_projection(src_seq)
        x = ops.add_pos_dropout(x, self.position_enc.weight, self._rng.make(self.p, self._site_in, dev, self.training))
        mask = get_attn_padding_mask(src_pad_mask, src_pad_mask) + get_attn_subsequent_mask(src_pad_mask, *self.sub)
        attns = []
        for layer in self.layer_stack:
            layer.slf_attn.return_attn = return_attns
            x, a = layer(x, slf_attn_mask=mask)
            attns.append(a)
        x = ops.add_pos_dropout(x, self.trans_pos_enc.weight, self._rng.make(self.p, self._site_out, dev, self.training))
        return (x, attns) if return_attns else (x,)


class EncoderTest(nn.Module):
    """The wired encoder (T/Models.py:127-166): splice(+-2) -> frozen LDA -> Linear -> dropout -> TDNN stack -> + pos ->
    dropout.  Returns a tensor (not a tuple), like the reference."""

    def __init__(self, lda_mat, n_src_dim, encoder_max_len, d_model=256, dropout=0.1, contexts=[[0]], rng=None):
        super().__init__()
        self.d_model = d_model
        self.p = float(dropout)
        self._rng = rng or _rng.GLOBAL
        self.trans_pos_enc = _frozen_table(encoder_max_len, d_model)
        self.concat = ConcatLayer(LDA_CONCAT_INDEX)
        self.lda_layer = LDALayer(lda_mat)
        self.src_projection = Linear(n_src_dim * len(LDA_CONCAT_INDEX), d_model, bias=False)
        self._site_src = self._rng.site("enc.src")
        self.tdnn_stack = nn.ModuleList([
            TDNNLayer(d_model, d_model, ctx, dropout=dropout, rng=self._rng, site="enc.tdnn.%d" % i)
            for i, ctx in enumerate(contexts)])
        self._site_out = self._rng.site("enc.out")

    def forward(self, src_seq, src_pad_mask=None, fold=1, cmvn=0):
        """`fold`/`cmvn` are extensions: when given, frame folding and per-utterance CMVN run inside the same front-end
        kernel as the splice (the reference folds with a separate view+copy and leaves CMVN to Kaldi's apply-cmvn)."""
        dev = src_seq.device
        lengths = None
        if cmvn:
            assert src_pad_mask is not None, "CMVN needs the pad mask to know the utterance lengths"
            lengths = src_pad_mask.to(torch.int32).sum(dim=1, dtype=torch.int32)
        bf16 = ops.compute_mode() == "bf16"     # activations of the TDNN stack live in HBM as bf16, GEMMs on tcgen05
        x = ops.frontend(src_seq, lengths, fold, self.concat.index, cmvn,
                         out_dtype=torch.bfloat16 if bf16 else torch.float32)
        x = self.lda_layer(x)
        # (single_consumer: the dropped-out projection feeds only TDNN layer 0, whose data-gradient GEMM then applies the
        #  mask of the backward pass in its epilogue -- likewise for the two dropouts further down)
        x = self.src_projection(x, drop=self._rng.make(self.p, self._site_src, dev, self.training), single_consumer=True)
        for layer in self.tdnn_stack:
            x = layer(x)
        return ops.add_pos_dropout(x, self.trans_pos_enc.weight, self._rng.make(self.p, self._site_out, dev, self.training),
                                   single_consumer=True)


class Decoder(nn.Module):
    """T/Models.py:169-231.  Returns a 1-tuple (logits,) like the reference."""

    def __init__(self, n_tgt_vocab, decoder_max_len, n_layers=2, n_head=3, sub_sequence=(-1, 1), d_k=64, d_v=64,
                 en_d_model=256, de_d_model=128, d_inner_hid=128, dropout=0.1, rng=None):
        super().__init__()
        self.sub = sub_sequence
        self.en_d_model, self.de_d_model = en_d_model, de_d_model
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.p = float(dropout)
        self._rng = rng or _rng.GLOBAL
        self.position_enc = _frozen_table(decoder_max_len, de_d_model)
        self.tgt_word_emb = nn.Embedding(n_tgt_vocab, de_d_model, padding_idx=constants.PAD)
        self.tgt_word_proj = Linear(de_d_model, n_tgt_vocab, bias=False)
        self._site_emb = self._rng.site("dec.emb")
        self.layer_stack = nn.ModuleList([
            DecoderLayer(de_d_model, d_inner_hid, n_head, d_k, d_v, dropout=dropout, rng=self._rng, site="dec.%d" % i)
            for i in range(n_layers)])
        self._site_out = self._rng.site("dec.out")
        self.enc_dec_projection = Linear(en_d_model, de_d_model, bias=False)

    def forward(self, tgt_seq, tgt_pad_mask, src_pad_mask, enc_output, return_attns=False):
        dev = enc_output.device
        # bf16 mode: the projected encoder memory stays bf16 -- its only consumers are the cross-attention K/V
        # projections (the largest decoder GEMMs: all T frames, every layer), which then run on the tensor cores too
        bf16 = ops.compute_mode() == "bf16" and tgt_seq.size(1) > 1
        if bf16 and enc_output.dtype != torch.bfloat16:
            enc_output = ops.cast(enc_output, torch.bfloat16)
        enc = self.enc_dec_projection(enc_output, out_fp32=not bf16)
        x = ops.embed_pos(tgt_seq, self.tgt_word_emb.weight, self.position_enc.weight,
                          self._rng.make(self.p, self._site_emb, dev, self.training), constants.PAD,
                          out_dtype=torch.bfloat16 if bf16 else torch.float32)
        slf_mask = get_attn_padding_mask(tgt_pad_mask, tgt_pad_mask) + get_attn_subsequent_mask(tgt_pad_mask, *self.sub)
        enc_mask = get_attn_padding_mask(tgt_pad_mask, src_pad_mask)
        slf_attns, enc_attns = [], []
        # bf16 path: the cross-attention k|v projections of ALL layers are one GEMM over the encoder memory (the largest
        # decoder GEMMs: every frame, every layer); layer l reads its column block
        enc_kv = ops.cross_kv_proj(enc, [(l.enc_attn.w_ks, l.enc_attn.w_vs) for l in self.layer_stack]) \
            if (bf16 and not return_attns) else None
        for li, layer in enumerate(self.layer_stack):
            layer.slf_attn.return_attn = layer.enc_attn.return_attn = return_attns
            layer.enc_attn.kv_override = enc_kv[li] if enc_kv is not None else None
            x, a1, a2 = layer(x, enc, slf_attn_mask=slf_mask, dec_enc_attn_mask=enc_mask)
            slf_attns.append(a1)
            enc_attns.append(a2)
        x = ops.add_pos_dropout(x, None, self._rng.make(self.p, self._site_out, dev, self.training), single_consumer=True)
        # fp32 logits for the loss.  bf16 path: tensor-core GEMM with fp32 output; V = 53 does not meet TMA's 16-byte row
        # pitch as a data-gradient operand, so the backward zero-pads the logit gradient to 56 columns (ops.linear_tc)
        logits = self.tgt_word_proj(x, out_fp32=True) if x.dtype == torch.bfloat16 else self.tgt_word_proj(x)
        return (logits, slf_attns, enc_attns) if return_attns else (logits,)


class Transformer(nn.Module):
    """Constructor and forward signature of T/Models.py:233-261.  Extra keyword-only knobs (all default to the
    reference behaviour): encoder_type ("tdnn" | "attention"), cmvn ("none" | "mean" | "meanvar"), seed (dropout)."""

    def __init__(self, n_src_dim, n_tgt_vocab, lda_mat, encoder_max_len, decoder_max_len, src_fold=1,
                 encoder_sub_sequence=(-100, 0), decoder_sub_sequence=(-20, 0), en_layers=2, de_layers=2, n_head=3,
                 en_d_model=256, de_d_model=128, d_k=64, d_v=64, en_dropout=0.2, de_dropout=0.2, tdnn_contexts=[[0]],
                 *, encoder_type="tdnn", cmvn="none", seed=0):
        super().__init__()
        self.src_fold = src_fold
        self.encoder_type = encoder_type
        self.cmvn = cmvn
        self.dropout_state = _rng.DropoutState(seed)
        self._operands = ops.OperandCache()      # bf16 GEMM operand copies of the weights, one refresh launch per forward
        if encoder_type == "tdnn":
            self.encoder_test = EncoderTest(lda_mat=lda_mat, n_src_dim=n_src_dim * src_fold,
                                            encoder_max_len=encoder_max_len, d_model=en_d_model, dropout=en_dropout,
                                            contexts=tdnn_contexts, rng=self.dropout_state)
        elif encoder_type == "attention":
            self.encoder = Encoder(n_src_dim=n_src_dim * src_fold, encoder_max_len=encoder_max_len,
                                   sub_sequence=encoder_sub_sequence, n_layers=en_layers, n_head=n_head, d_k=d_k, d_v=d_v,
                                   d_model=en_d_model, d_inner_hid=en_d_model, dropout=en_dropout, rng=self.dropout_state)
        else:
            raise ValueError("encoder_type must be 'tdnn' or 'attention'")
        # Reference quirk kept for weight interchange: T/Models.py:250-251 does NOT forward d_k/d_v to the Decoder, so
        # the decoder heads always use the Decoder defaults (64) whatever d_k/d_v the Transformer was given.
        self.decoder = Decoder(n_tgt_vocab=n_tgt_vocab, decoder_max_len=decoder_max_len, sub_sequence=decoder_sub_sequence,
                               n_layers=de_layers, n_head=n_head, en_d_model=en_d_model,
                               de_d_model=de_d_model, d_inner_hid=de_d_model, dropout=de_dropout, rng=self.dropout_state)

    @property
    def dropout_sites(self):
        return dict(self.dropout_state.sites)

    def encode(self, src_seq, src_pad_mask):
        """fold + encoder -> (enc_output, folded pad mask)."""
        fold = self.src_fold
        if self.encoder_type == "tdnn":
            mask = src_pad_mask if fold == 1 else src_pad_mask[:, fold - 1::fold].contiguous()
            return self.encoder_test(src_seq, src_pad_mask, fold=fold, cmvn=CMVN_MODES[self.cmvn]), mask
        seq, mask = fold_seq_and_mask(src_seq, src_pad_mask, fold)
        return self.encoder(seq, mask)[0], mask

    def forward(self, src_seq, src_pad_mask, tgt_seq, tgt_pad_mask):
        ops.reset_deferred()
        if ops.compute_mode() == "bf16" and tgt_seq.size(1) > 1:
            # tensor-core path: ONE launch refreshes every bf16 operand copy of the weights (they only change in the
            # optimiser step) and advances the dropout step counter; the ops below find their operands in the cache
            counter = self.dropout_state.step_tensor(src_seq.device) if self.training else None
            self._operands.refresh(counter)
            with self._operands.active():
                enc_output, mask = self.encode(src_seq, src_pad_mask)
                dec_output, *_ = self.decoder(tgt_seq, tgt_pad_mask, mask, enc_output)
            return dec_output
        if self.training:
            self.dropout_state.tick(src_seq.device)
        enc_output, mask = self.encode(src_seq, src_pad_mask)
        dec_output, *_ = self.decoder(tgt_seq, tgt_pad_mask, mask, enc_output)
        return dec_output
