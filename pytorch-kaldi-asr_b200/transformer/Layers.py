"""Layer composition (T/Layers.py): post-LN sub-layers chained; the attention maps are optional outputs."""
from __future__ import annotations

import torch.nn as nn

from .SubLayers import MultiHeadAttention, PositionwiseFeedForward


class EncoderLayer(nn.Module):
    def __init__(self, d_model, d_inner_hid, n_head, d_k, d_v, dropout=0.1, rng=None, site="enc.0"):
        super().__init__()
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout, rng=rng, site=site + ".slf")
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner_hid, dropout=dropout, rng=rng, site=site + ".ffn")

    def forward(self, enc_input, slf_attn_mask=None):
        out, attn = self.slf_attn(enc_input, enc_input, enc_input, attn_mask=slf_attn_mask)
        return self.pos_ffn(out), attn


class DecoderLayer(nn.Module):
    def __init__(self, d_model, d_inner_hid, n_head, d_k, d_v, dropout=0.1, rng=None, site="dec.0"):
        super().__init__()
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout, rng=rng, site=site + ".slf")
        self.enc_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout, rng=rng, site=site + ".enc")
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner_hid, dropout=dropout, rng=rng, site=site + ".ffn")

    def forward(self, dec_input, enc_output, slf_attn_mask=None, dec_enc_attn_mask=None):
        out, slf = self.slf_attn(dec_input, dec_input, dec_input, attn_mask=slf_attn_mask)
        out, enc = self.enc_attn(out, enc_output, enc_output, attn_mask=dec_enc_attn_mask)
        return self.pos_ffn(out), slf, enc
