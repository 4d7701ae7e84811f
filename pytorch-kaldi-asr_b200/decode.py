"""Batched beam-search decoding with the reference's entry point `translate_batch(model, batch, opt, model_options)`
(L/decode.py:22-107), re-designed for the GPU:

  reference (per step)                                   here (per step)
  ---------------------------------------------------    -----------------------------------------------------------
  re-run the WHOLE decoder over every full prefix         one new token per live hypothesis; self-attention reads a
  (O(step^2) work, L/decode.py:85)                        per-edge KV cache along the lattice back-pointers
  index_select a [n_live, T, De] copy of the encoder      cross-attention K/V projected ONCE per utterance and shared by
  output and re-project it to K/V in every layer          its hypotheses (the beam is the query axis of one attention call)
  D2H of the log-probs, numpy Lattice.advance on host,    device-resident lattice + warp-level top-k kernel; the step is a
  python lists -> LongTensor -> H2D (:60-98)              CUDA graph replay; the host only polls a "not done" counter

Semantics kept exactly (Appendix A.14, section 7 "hard parts" of SURVEY.md): float64 hypothesis scores = running sum of
fp32 log-probs; finished hypotheses stay in the beam and compete; no length normalisation; the first step expands
only BOS; LayerNormalization is the identity at step 1 (sequence length 1, T/Modules.py:43-44) while the cached K/V of
position 0 come from a LayerNorm-applied pass (which is what the reference recomputes from step 2 on); ties are broken
towards the lowest flat candidate index (the reference's np.argsort order for exact ties is unspecified).

Only a causal decoder band (end == 0) can be cached: with end > 0 an earlier position attends to later tokens, so its
keys/values change whenever the hypothesis grows.  Such bands -- which the reference decodes like any other, because it
never caches (T/Models.py:38-49, L/decode.py:85) -- take the NO-CACHE path (`BeamDecoder(..., use_cache=False)`,
chosen automatically): every step re-runs `model.decoder` over the full prefixes of all beam slots through the same
fp32 kernels as training/evaluation, and the device lattice / top-k kernel advances as in the cached path.  It is
O(step^2) like the reference and exists for exactness, not speed; `opt.use_cache = False` forces it for a causal band
(the two paths are tested against each other and against the reference's goldens).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List

import numpy as np
import torch

from . import _lib as L
from . import ops
from .utils import constants


class BeamDecoder:
    """Decoding state and step graph for one (n_utt, T, beam, max_len) shape."""

    def __init__(self, model, n_utt: int, T: int, beam: int, max_len: int, force_full_length: bool = False,
                 use_graph: bool = True, poll_every: int = 8, use_cache: bool = True):
        dec = model.decoder
        self.cached = bool(use_cache) and int(dec.sub[1]) == 0     # a non-causal band cannot be cached (module docstring)
        if max_len > dec.position_enc.weight.shape[0]:
            raise RuntimeError("max_token_seq_len %d exceeds decoder_max_len %d" % (max_len, dec.position_enc.weight.shape[0]))
        self.model, self.dec = model, dec
        self.n_utt, self.T, self.beam, self.max_len = n_utt, T, beam, max_len
        self.force, self.use_graph, self.poll_every = force_full_length, use_graph, poll_every
        self.dev = dev = dec.tgt_word_emb.weight.device
        self.H, self.dk, self.D = dec.n_head, dec.d_k, dec.de_d_model
        self.HD = self.H * self.dk
        self.V = dec.tgt_word_emb.weight.shape[0]
        self.window = -int(dec.sub[0]) + 1
        self.E = 1 + beam * max_len
        self.n_layers = len(dec.layer_stack)
        i32 = dict(device=dev, dtype=torch.int32)
        self.edge_prev = torch.empty(n_utt, self.E, **i32)
        self.edge_word = torch.empty(n_utt, self.E, **i32)
        self.edge_depth = torch.empty(n_utt, self.E, **i32)
        self.edge_weight = torch.empty(n_utt, self.E, device=dev, dtype=torch.float64)
        self.n_edges = torch.empty(n_utt, **i32)
        self.beam_edges = torch.empty(n_utt, beam, **i32)
        self.beam_count = torch.empty(n_utt, **i32)
        self.slot_edge = torch.empty(n_utt, beam, **i32)
        self.slot_active = torch.empty(n_utt, beam, **i32)
        self.curr_length = torch.empty(n_utt, **i32)
        self.done = torch.empty(n_utt, **i32)
        self.n_not_done = torch.empty(1, **i32)
        f32 = dict(device=dev, dtype=torch.float32)
        n_c = self.n_layers if self.cached else 0                  # the no-cache path keeps none of these
        self.kcache = [torch.empty(n_utt, self.E, self.HD, **f32) for _ in range(n_c)]
        self.vcache = [torch.empty(n_utt, self.E, self.HD, **f32) for _ in range(n_c)]
        self.enc_kv = [torch.empty(n_utt, T, 2 * self.HD, **f32) for _ in range(n_c)]
        self.enc_rep = self.mask_rep = None                        # no-cache path: encoder memory / mask per beam slot
        self.src_mask = torch.empty(n_utt, T, device=dev, dtype=torch.uint8)
        self.logits = torch.empty(n_utt * beam, self.V, **f32)
        self.desc = L.BeamDesc()
        self.desc.n_utt, self.desc.beam, self.desc.V = n_utt, beam, self.V
        self.desc.max_edges, self.desc.max_len = self.E, max_len
        self.desc.eos, self.desc.force_full_length = constants.EOS, int(force_full_length)
        self.graph = None
        self.steps_run = 0
        # The weights are frozen while decoding: pack the per-head projection tensors w_p[h, d, j] (T/SubLayers.py:29-31)
        # once into nn.Linear layout [(p,h,j), d], so q|k|v of a step is ONE GEMM instead of three head-batched ones.
        # (re-packed into the same buffers on every reset(), so weight updates between batches are picked up and the
        # captured step graph keeps pointing at valid storage)
        self.w_qkv = [torch.empty(3 * self.HD, self.D, **f32) for _ in range(n_c)]
        self.w_q_cross = [torch.empty(self.HD, self.D, **f32) for _ in range(n_c)]

    # ------------------------------------------------------------------------------------------------ state
    def reset(self, enc_output: torch.Tensor, src_pad_mask: torch.Tensor):
        """Start a new batch: lattice = {BOS}, cross-attention K/V of every layer from the encoder output (once)."""
        assert enc_output.shape[0] == self.n_utt and enc_output.shape[1] == self.T
        self.edge_prev.fill_(-1)
        self.edge_word.zero_()
        self.edge_word[:, 0] = constants.BOS
        self.edge_depth.zero_()
        self.edge_weight.zero_()
        self.n_edges.fill_(1)
        self.beam_edges.zero_()
        self.beam_count.fill_(1)
        self.slot_edge.zero_()
        self.slot_active.zero_()
        self.slot_active[:, 0] = 1
        self.curr_length.zero_()
        self.done.zero_()
        self.n_not_done.fill_(self.n_utt)
        self.src_mask.copy_(src_pad_mask.to(torch.uint8))
        self.steps_run = 0
        if not self.cached:
            # like the reference (L/decode.py:64-75) every beam slot carries its own copy of the encoder memory
            self.enc_rep = enc_output.to(torch.float32).repeat_interleave(self.beam, 0).contiguous()
            self.mask_rep = self.src_mask.repeat_interleave(self.beam, 0).contiguous()
            return
        with torch.no_grad():
            pack = lambda *ws: torch.cat([w.detach().permute(0, 2, 1).reshape(-1, w.shape[1]) for w in ws])
            for l, layer in enumerate(self.dec.layer_stack):
                self.w_qkv[l].copy_(pack(layer.slf_attn.w_qs, layer.slf_attn.w_ks, layer.slf_attn.w_vs))
                self.w_q_cross[l].copy_(pack(layer.enc_attn.w_qs))
            enc = self.dec.enc_dec_projection(enc_output, out_fp32=True)                      # [n_utt, T, Dd], once (T/Models.py:199)
            for l, layer in enumerate(self.dec.layer_stack):
                self.enc_kv[l].copy_(ops.head_proj(enc, layer.enc_attn.w_ks, layer.enc_attn.w_vs))
        self.steps_run = 0

    # ------------------------------------------------------------------------------------------------ one token
    def _decoder_token(self, skip_ln: bool, write_cache: bool):
        """Run the decoder stack on the newest token of every slot -> self.logits.  `skip_ln`: the reference's
        LayerNormalization is the identity when the decoder input has length 1 (first step)."""
        lib, st = L.lib(), L.stream_ptr
        dec, n, K, HD, D = self.dec, self.n_utt, self.beam, self.HD, self.D
        scale = 1.0 / math.sqrt(D)                                             # temper = sqrt(d_model), T/Modules.py:72
        x = torch.empty(n, K, D, device=self.dev, dtype=torch.float32)
        L.check(lib.pka_beam_embed(L.ptr(dec.tgt_word_emb.weight), L.ptr(dec.position_enc.weight), L.ptr(self.edge_word),
                                   L.ptr(self.edge_depth), L.ptr(self.slot_edge), L.ptr(self.slot_active), L.ptr(x), n, K,
                                   self.E, D, st()), "beam_embed")
        for l, layer in enumerate(dec.layer_stack):
            sa, ca, ff = layer.slf_attn, layer.enc_attn, layer.pos_ffn
            # --- self-attention over the lattice ancestors
            qkv = ops.linear(x, self.w_qkv[l])                                 # [n, K, 3*HD]
            ctx = torch.empty(n, K, HD, device=self.dev, dtype=torch.float32)
            qp = qkv.data_ptr()
            L.check(lib.pka_tree_attn(C.c_void_p(qp), C.c_void_p(qp + 4 * HD), C.c_void_p(qp + 8 * HD), 3 * HD,
                                      L.ptr(self.kcache[l]), L.ptr(self.vcache[l]), L.ptr(self.edge_prev),
                                      L.ptr(self.slot_edge), L.ptr(self.slot_active), L.ptr(ctx), n, K, self.E, self.H,
                                      self.dk, self.window, C.c_float(scale), st()), "tree_attn")
            if write_cache:
                L.check(lib.pka_kv_append(C.c_void_p(qp + 4 * HD), C.c_void_p(qp + 8 * HD), 3 * HD, L.ptr(self.kcache[l]),
                                          L.ptr(self.vcache[l]), L.ptr(self.slot_edge), L.ptr(self.slot_active), n, K,
                                          self.E, HD, st()), "kv_append")
            x = self._post(sa.proj, ctx, x, sa.layer_norm, skip_ln)
            # --- cross-attention: the beam is the query axis, K/V shared per utterance
            q = ops.linear(x, self.w_q_cross[l])
            ctx, _ = ops.attention(q, self.enc_kv[l], self.src_mask, self.H, self.dk, None, scale, None)
            x = self._post(ca.proj, ctx, x, ca.layer_norm, skip_ln)
            # --- position-wise FFN
            h = ops.linear(x, ff.w_1.weight, ff.w_1.bias, relu=True)
            if skip_ln:
                x = ops.linear(h, ff.w_2.weight, ff.w_2.bias, residual=x)
            else:
                x = ops.add_layer_norm(ops.linear(h, ff.w_2.weight, ff.w_2.bias), x, ff.layer_norm.a_2, ff.layer_norm.b_2,
                                       ff.layer_norm.eps)
        logits = ops.linear(x, dec.tgt_word_proj.linear.weight, None)
        self.logits.copy_(logits.view(n * K, self.V))

    @staticmethod
    def _post(proj, ctx, residual, ln, skip_ln):
        if skip_ln:
            return proj(ctx, residual=residual)
        return ops.add_layer_norm(proj(ctx), residual, ln.a_2, ln.b_2, ln.eps)

    def _advance(self):
        L.check(L.lib().pka_beam_advance(C.byref(self.desc), L.ptr(self.logits), L.ptr(self.edge_prev), L.ptr(self.edge_word),
                                         L.ptr(self.edge_depth), L.ptr(self.edge_weight), L.ptr(self.n_edges),
                                         L.ptr(self.beam_edges), L.ptr(self.beam_count), L.ptr(self.slot_edge),
                                         L.ptr(self.slot_active), L.ptr(self.curr_length), L.ptr(self.done),
                                         L.ptr(self.n_not_done), L.stream_ptr()), "beam_advance")

    def _step_body(self):
        self._decoder_token(skip_ln=False, write_cache=True)
        self._advance()

    # ------------------------------------------------------------------------------------------------ no-cache path
    def _slot_prefixes(self, length: int) -> torch.Tensor:
        """Token prefixes int64[n_utt*beam, length] of the beam slots, read off the device lattice along the
        back-pointers.  Every live hypothesis has exactly `length` tokens; an idle slot points at the BOS edge and gets
        an all-BOS row (its logits are ignored by `beam_advance`)."""
        cur = self.slot_edge.long()
        prev, word = self.edge_prev.long(), self.edge_word.long()
        toks = torch.full((self.n_utt, self.beam, length), constants.BOS, device=self.dev, dtype=torch.int64)
        for pos in range(length - 1, -1, -1):
            alive = cur >= 0
            safe = cur.clamp(min=0)
            toks[:, :, pos] = torch.where(alive, word.gather(1, safe), toks[:, :, pos])
            cur = torch.where(alive, prev.gather(1, safe), cur)
        return toks.view(self.n_utt * self.beam, length)

    def _run_uncached(self):
        """The reference's own scheme (L/decode.py:54-98): per step, the whole decoder over the full prefixes.  The
        decoder call is the model's ordinary forward (fp32 kernels; LayerNormalization is the identity at length 1 by
        itself, T/Modules.py:43-44); the lattice update stays on the device."""
        prev_mode = ops.compute_mode()
        ops.set_compute_mode("fp32")
        try:
            steps = 0
            while steps < self.max_len:
                tgt = self._slot_prefixes(steps + 1)
                ones = torch.ones(tgt.shape, device=self.dev, dtype=torch.uint8)      # only the band restricts
                logits = self.dec(tgt, ones, self.mask_rep, self.enc_rep)[0]            # [n*K, length, V]
                self.logits.copy_(logits[:, -1, :])
                self._advance()
                steps += 1
                if steps % self.poll_every == 0 and int(self.n_not_done.item()) == 0:
                    break
        finally:
            ops.set_compute_mode(prev_mode)
        self.steps_run = steps
        return steps

    @torch.no_grad()
    def run(self):
        """Decode until every lattice is done or max_len steps were taken.  Returns the number of steps run."""
        if not self.cached:
            return self._run_uncached()
        # step 1: position 0 twice -- LN-applied pass fills the caches of the BOS edge, LN-skipped pass gives the logits
        self._decoder_token(skip_ln=False, write_cache=True)
        self._decoder_token(skip_ln=True, write_cache=False)
        self._advance()
        steps = 1
        if self.max_len > 1:
            if self.use_graph and self.graph is None:
                # capture the steady-state step once; lattice state is saved/restored around warm-up and capture
                snap = [t.clone() for t in self._state()]
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._step_body()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                for t, s in zip(self._state(), snap):
                    t.copy_(s)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._step_body()
                for t, s in zip(self._state(), snap):
                    t.copy_(s)
            while steps < self.max_len:
                if self.graph is not None:
                    self.graph.replay()
                else:
                    self._step_body()
                steps += 1
                if steps % self.poll_every == 0 and int(self.n_not_done.item()) == 0:
                    break
        self.steps_run = steps
        return steps

    def _state(self):
        return [self.edge_prev, self.edge_word, self.edge_depth, self.edge_weight, self.n_edges, self.beam_edges,
                self.beam_count, self.slot_edge, self.slot_active, self.curr_length, self.done, self.n_not_done]

    # ------------------------------------------------------------------------------------------------ read-out
    def results(self, nbest: int):
        """-> (hyps[n_utt][<=nbest][tokens], weights[n_utt][beam entries]) like Lattice.get_results('all') per utterance
        (T/Lattice.py:93-107).  One D2H of the lattice arrays, then a vectorised back-pointer walk."""
        prev = self.edge_prev.cpu().numpy()
        word = self.edge_word.cpu().numpy()
        depth = self.edge_depth.cpu().numpy()
        weight = self.edge_weight.cpu().numpy()
        beam_edges = self.beam_edges.cpu().numpy()
        beam_count = self.beam_count.cpu().numpy()
        n, K = beam_edges.shape
        u_idx = np.arange(n)[:, None].repeat(K, 1)
        cur = beam_edges.copy()
        max_depth = int(depth[u_idx, cur].max())
        toks = np.zeros((n, K, max_depth + 1), dtype=np.int64)
        lens = depth[u_idx, cur] + 1
        for _ in range(max_depth + 1):
            alive = cur >= 0
            safe = np.where(alive, cur, 0)
            d = depth[u_idx, safe]
            toks[u_idx[alive], np.nonzero(alive)[1], d[alive]] = word[u_idx, safe][alive]
            cur = np.where(alive, prev[u_idx, safe], -1)
        hyps, weights = [], []
        for u in range(n):
            c = int(beam_count[u])
            hyps.append([toks[u, k, :lens[u, k]].tolist() for k in range(min(c, nbest))])
            weights.append([float(weight[u, beam_edges[u, k]]) for k in range(c)])
        return hyps, weights


_DECODERS = {}


def translate_batch(model, batch, opt, model_options=None):
    """Reference signature (L/decode.py:22).  `batch` = (keys, src f32[B,T,F], src_pad_mask u8[B,T], tgt, tgt_mask);
    `opt` carries beam_size, max_token_seq_len, nbest (use_gpu is accepted and ignored: this path is GPU only).
    Optional extras on `opt`: force_full_length (bool), use_graph (bool), use_cache (bool; False = the reference's
    no-cache scheme, which is also what a non-causal decoder band gets)."""
    model.eval()
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("translate_batch: the B200 path has no CPU mode; move the model to a CUDA device first")
    src = torch.as_tensor(np.ascontiguousarray(batch[1]), dtype=torch.float32).to(dev, non_blocking=True) \
        if not torch.is_tensor(batch[1]) else batch[1].to(dev, torch.float32, non_blocking=True)
    mask = torch.as_tensor(np.ascontiguousarray(batch[2]), dtype=torch.uint8).to(dev, non_blocking=True) \
        if not torch.is_tensor(batch[2]) else batch[2].to(dev, torch.uint8, non_blocking=True)
    with torch.no_grad():
        enc_output, fmask = model.encode(src, mask)
    # the captured step graph holds raw pointers to the decoder weights: the storage fingerprint makes a model whose
    # parameters moved (FusedAdam arena, .to(), dtype change) build a fresh decoder instead of replaying stale addresses
    fingerprint = hash(tuple(p.data_ptr() for p in model.decoder.parameters()))
    use_cache = bool(getattr(opt, "use_cache", True)) and int(model.decoder.sub[1]) == 0
    n_utt, T = enc_output.shape[0], enc_output.shape[1]
    # the no-cache path replicates the encoder memory per beam slot: bound it to ~1 GiB by decoding utterance chunks
    chunk = n_utt if use_cache else max(1, min(n_utt, (1 << 28) // max(1, opt.beam_size * T * enc_output.shape[2])))
    hyps, weights = [], []
    for u0 in range(0, n_utt, chunk):
        n = min(chunk, n_utt - u0)
        key = (id(model), n, T, opt.beam_size, opt.max_token_seq_len, bool(getattr(opt, "force_full_length", False)),
               bool(getattr(opt, "use_graph", True)), fingerprint, use_cache)
        bd = _DECODERS.get(key)
        if bd is None:
            if len(_DECODERS) > 8:
                _DECODERS.clear()
            bd = _DECODERS[key] = BeamDecoder(model, n, T, key[3], key[4], key[5], key[6], use_cache=use_cache)
        bd.reset(enc_output[u0:u0 + n], fmask[u0:u0 + n])
        bd.run()
        h, w = bd.results(opt.nbest)
        hyps += h
        weights += w
    return hyps, weights
