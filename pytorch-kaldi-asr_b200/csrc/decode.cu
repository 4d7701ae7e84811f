// Beam-search decoding kernels: a device-resident lattice per utterance, a per-edge KV cache with tree attention,
// and a warp-level top-k.  Together they replace the body of the reference's decoding loop -- full-prefix decoder
// re-runs, per-step D2H of the log-probs and numpy Lattice.advance on the host (L/decode.py:54-98,
// T/Lattice.py:35-81) -- with O(1) work per new token and no per-step host round trip.
//
// Lattice layout (all arrays on the device, one row per utterance, E = max_edges = 1 + beam*max_len):
//   edge_prev/edge_word/edge_depth int32[n_utt,E], edge_weight f64[n_utt,E]   every edge ever created (edge 0 = BOS)
//   n_edges int32[n_utt]
//   beam_edges int32[n_utt,beam], beam_count int32[n_utt]    the current beam, best first (finished edges stay in it)
//   slot_edge int32[n_utt,beam], slot_active int32[n_utt,beam]   live hypotheses in beam order = rows of the next
//                                                                decoder step (row = u*beam + slot)
//   done int32[n_utt], curr_length int32[n_utt], n_not_done int32[1]
// KV cache: kcache/vcache f32[n_utt, E, H*dk] per decoder layer, written once per edge; the self-attention of a
// hypothesis walks its back-pointers (at most `window` = -band_start+1 keys: itself + ancestors).
#include "common.cuh"
#include <math_constants.h>

namespace pka {

// ------------------------------------------------------------------------------------------------ embedding of new tokens
__global__ void beam_embed_kernel(const float* __restrict__ emb, const float* __restrict__ pos,
                                  const int* __restrict__ edge_word, const int* __restrict__ edge_depth,
                                  const int* __restrict__ slot_edge, const int* __restrict__ slot_active,
                                  float* __restrict__ out, int n_slots, int beam, int max_edges, int D) {
  pdl_wait();
  const int d4 = D >> 2;
  const long long total = (long long)n_slots * d4;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % d4) * 4;
    const int s = (int)(e / d4);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slot_active[s]) {
      const int u = s / beam;
      const long long ed = (long long)u * max_edges + slot_edge[s];
      const float4 ev = *reinterpret_cast<const float4*>(emb + (long long)edge_word[ed] * D + c);
      const float4 pv = *reinterpret_cast<const float4*>(pos + (long long)edge_depth[ed] * D + c);
      o = make_float4(ev.x + pv.x, ev.y + pv.y, ev.z + pv.z, ev.w + pv.w);
    }
    *reinterpret_cast<float4*>(out + (long long)s * D + c) = o;
  }
}

// ------------------------------------------------------------------------------------------------ KV cache append
__global__ void kv_append_kernel(const float* __restrict__ k_new, const float* __restrict__ v_new, int ld,
                                 float* __restrict__ kcache, float* __restrict__ vcache,
                                 const int* __restrict__ slot_edge, const int* __restrict__ slot_active, int n_slots,
                                 int beam, int max_edges, int HD) {
  pdl_wait();
  const int d4 = HD >> 2;
  const long long total = (long long)n_slots * d4;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % d4) * 4;
    const int s = (int)(e / d4);
    if (!slot_active[s]) continue;
    const int u = s / beam;
    const long long dst = ((long long)u * max_edges + slot_edge[s]) * HD + c;
    *reinterpret_cast<float4*>(kcache + dst) = *reinterpret_cast<const float4*>(k_new + (long long)s * ld + c);
    *reinterpret_cast<float4*>(vcache + dst) = *reinterpret_cast<const float4*>(v_new + (long long)s * ld + c);
  }
}

// ------------------------------------------------------------------------------------------------ tree attention
// one warp per (slot, head).  Keys: the fresh k/v of the slot's own token (from the packed qkv buffer) plus up to
// window-1 ancestors read from the per-edge cache along edge_prev.
constexpr int kTreeMaxWindow = 128;
template <int D>
__global__ void __launch_bounds__(128)
tree_attn_kernel(const float* __restrict__ q, const float* __restrict__ k_self, const float* __restrict__ v_self, int ld,
                 const float* __restrict__ kcache, const float* __restrict__ vcache, const int* __restrict__ edge_prev,
                 const int* __restrict__ slot_edge, const int* __restrict__ slot_active, float* __restrict__ out,
                 int n_slots, int beam, int max_edges, int H, int window, float scale) {
  pdl_wait();
  constexpr int R = D / 32 > 0 ? D / 32 : 1;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_slots * H) return;
  const int s = warp / H, h = warp % H;
  float* orow = out + (long long)s * (H * D) + h * D;
  if (!slot_active[s]) {
#pragma unroll
    for (int r = 0; r < R; ++r) { const int d = lane + 32 * r; if (d < D) orow[d] = 0.f; }
    return;
  }
  const int u = s / beam;
  const int* prev_u = edge_prev + (long long)u * max_edges;
  float qv[R], acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int d = lane + 32 * r;
    qv[r] = d < D ? q[(long long)s * ld + h * D + d] : 0.f;
    acc[r] = 0.f;
  }
  float m_run = -CUDART_INF_F, l_run = 0.f;
  int e = slot_edge[s];
  for (int w = 0; w < window && e >= 0; ++w) {
    const float* kr = (w == 0) ? k_self + (long long)s * ld + h * D : kcache + ((long long)u * max_edges + e) * (H * D) + h * D;
    const float* vr = (w == 0) ? v_self + (long long)s * ld + h * D : vcache + ((long long)u * max_edges + e) * (H * D) + h * D;
    float dot = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) { const int d = lane + 32 * r; if (d < D) dot = fmaf(qv[r], kr[d], dot); }
    const float sc = warp_sum(dot) * scale;
    const float m_new = fmaxf(m_run, sc);
    const float corr = __expf(m_run - m_new);            // exp(-inf) = 0 on the first key
    const float pw = __expf(sc - m_new);
    l_run = l_run * corr + pw;
#pragma unroll
    for (int r = 0; r < R; ++r) { const int d = lane + 32 * r; if (d < D) acc[r] = acc[r] * corr + pw * vr[d]; }
    m_run = m_new;
    e = prev_u[e];
  }
  const float inv = 1.f / l_run;
#pragma unroll
  for (int r = 0; r < R; ++r) { const int d = lane + 32 * r; if (d < D) orow[d] = acc[r] * inv; }
}

// ------------------------------------------------------------------------------------------------ lattice advance + top-k
// One CTA (4 warps) per utterance.  Candidate i < n_act*V is (active hypothesis i/V in beam order, word i%V) with
// score weight[parent] + log_softmax(logits)[word] in fp64 (fp32 log-prob promoted, like numpy in T/Lattice.py:45);
// then the finished hypotheses in beam order.  `beam` rounds of arg-max (warp shuffles, then across warps), ties to the
// lowest candidate index.  Thread 0 then rewrites the lattice exactly like Lattice.advance.
__global__ void __launch_bounds__(128)
beam_advance_kernel(const pka_beam_desc d, const float* __restrict__ logits, int* __restrict__ edge_prev,
                    int* __restrict__ edge_word, int* __restrict__ edge_depth, double* __restrict__ edge_weight,
                    int* __restrict__ n_edges, int* __restrict__ beam_edges, int* __restrict__ beam_count,
                    int* __restrict__ slot_edge, int* __restrict__ slot_active, int* __restrict__ curr_length,
                    int* __restrict__ done, int* __restrict__ n_not_done) {
  pdl_wait();
  extern __shared__ double cand[];                       // [beam*V + beam]
  __shared__ float lse_s[64];
  __shared__ double wbest[4];
  __shared__ int ibest[4];
  __shared__ int pick_idx[64];
  __shared__ double pick_val[64];
  const int u = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (done[u]) return;
  const int V = d.V, beam = d.beam, E = d.max_edges;
  int* prev_u = edge_prev + (long long)u * E;
  int* word_u = edge_word + (long long)u * E;
  int* depth_u = edge_depth + (long long)u * E;
  double* wt_u = edge_weight + (long long)u * E;
  int* beam_u = beam_edges + (long long)u * beam;
  int* sedge_u = slot_edge + (long long)u * beam;
  int* sact_u = slot_active + (long long)u * beam;
  int n_act = 0;
  for (int k = 0; k < beam; ++k) n_act += sact_u[k];
  const int n_beam = beam_count[u];
  const int n_fin = n_beam - n_act;
  // log-sum-exp per active row (fp32, like nn.LogSoftmax on the fp32 logits)
  for (int a = warp; a < n_act; a += 4) {
    const float* row = logits + ((long long)u * beam + a) * V;
    float mx = -CUDART_INF_F;
    for (int c = lane; c < V; c += 32) mx = fmaxf(mx, row[c]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < V; c += 32) se += expf(row[c] - mx);
    se = warp_sum(se);
    if (lane == 0) lse_s[a] = d.inputs_are_logprobs ? 0.f : mx + logf(se);
  }
  __syncthreads();
  const int n_live_cand = n_act * V;
  const int n_cand = n_live_cand + n_fin;
  for (int i = tid; i < n_live_cand; i += 128) {
    const int a = i / V, w = i % V;
    float lp = logits[((long long)u * beam + a) * V + w] - lse_s[a];
    if (d.force_full_length && w == d.eos) lp = -1e30f;
    cand[i] = wt_u[sedge_u[a]] + (double)lp;
  }
  if (tid == 0) {                                        // finished hypotheses, in beam order
    int f = 0;
    for (int k = 0; k < n_beam; ++k) {
      const int e = beam_u[k];
      if (word_u[e] == d.eos) cand[n_live_cand + (f++)] = wt_u[e];
    }
  }
  __syncthreads();
  const int n_pick = n_cand < beam ? n_cand : beam;
  for (int r = 0; r < n_pick; ++r) {
    double bv = -CUDART_INF;
    int bi = 0x7fffffff;
    for (int i = tid; i < n_cand; i += 128) {
      const double v = cand[i];
      if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { wbest[warp] = bv; ibest[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 4; ++w)
        if (wbest[w] > bv || (wbest[w] == bv && ibest[w] < bi)) { bv = wbest[w]; bi = ibest[w]; }
      pick_idx[r] = bi; pick_val[r] = bv;
      if (bi != 0x7fffffff) cand[bi] = -CUDART_INF;      // remove from the next rounds
    }
    __syncthreads();
  }
  if (tid == 0) {
    // finished edges in old-beam order (to map candidate index -> edge id)
    int fin_edges[64];
    int f = 0;
    for (int k = 0; k < n_beam; ++k) if (word_u[beam_u[k]] == d.eos) fin_edges[f++] = beam_u[k];
    int old_slots[64];
    for (int k = 0; k < beam; ++k) old_slots[k] = sedge_u[k];
    int ne = n_edges[u], n_new_act = 0;
    for (int r = 0; r < n_pick; ++r) {
      const int i = pick_idx[r];
      int e;
      if (i < n_live_cand) {
        const int parent = old_slots[i / V];
        e = ne++;
        prev_u[e] = parent; word_u[e] = i % V; wt_u[e] = pick_val[r]; depth_u[e] = depth_u[parent] + 1;
      } else {
        e = fin_edges[i - n_live_cand];
      }
      beam_u[r] = e;
      if (word_u[e] != d.eos) { sedge_u[n_new_act] = e; sact_u[n_new_act] = 1; ++n_new_act; }
    }
    for (int k = n_new_act; k < beam; ++k) { sact_u[k] = 0; sedge_u[k] = 0; }
    n_edges[u] = ne;
    beam_count[u] = n_pick;
    const int len = curr_length[u] + 1;
    curr_length[u] = len;
    if (n_new_act == 0 || len > d.max_len) {
      done[u] = 1;
      for (int k = 0; k < beam; ++k) sact_u[k] = 0;
      atomicSub(n_not_done, 1);
    }
  }
}

}  // namespace pka

using namespace pka;

extern "C" int pka_beam_embed(const float* emb, const float* pos, const int32_t* edge_word, const int32_t* edge_depth,
                              const int32_t* slot_edge, const int32_t* slot_active, float* out, int n_utt, int beam,
                              int max_edges, int D, void* stream) {
  PKA_REQUIRE(emb && pos && edge_word && edge_depth && slot_edge && slot_active && out, PKA_EINVAL, "beam_embed: null pointer");
  PKA_REQUIRE(n_utt > 0 && beam > 0 && D % 4 == 0, PKA_EUNSUPPORTED, "beam_embed: n_utt=%d beam=%d D=%d", n_utt, beam, D);
  const long long total = (long long)n_utt * beam * (D / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  launch_k(beam_embed_kernel, blocks, 256, 0, as_stream(stream), emb, pos, edge_word, edge_depth, slot_edge, slot_active, out, n_utt * beam, beam, max_edges, D);
  return check_launch("beam_embed");
}

extern "C" int pka_kv_append(const float* k_new, const float* v_new, int ld, float* kcache, float* vcache,
                             const int32_t* slot_edge, const int32_t* slot_active, int n_utt, int beam, int max_edges,
                             int HD, void* stream) {
  PKA_REQUIRE(k_new && v_new && kcache && vcache && slot_edge && slot_active, PKA_EINVAL, "kv_append: null pointer");
  PKA_REQUIRE(HD % 4 == 0 && ld % 4 == 0 && aligned16(k_new) && aligned16(v_new), PKA_EALIGN, "kv_append: HD/ld must be multiples of 4 and pointers 16-byte aligned");
  const long long total = (long long)n_utt * beam * (HD / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  launch_k(kv_append_kernel, blocks, 256, 0, as_stream(stream), k_new, v_new, ld, kcache, vcache, slot_edge, slot_active, n_utt * beam, beam, max_edges, HD);
  return check_launch("kv_append");
}

extern "C" int pka_tree_attn(const float* q, const float* k_self, const float* v_self, int ld, const float* kcache,
                             const float* vcache, const int32_t* edge_prev, const int32_t* slot_edge,
                             const int32_t* slot_active, float* out, int n_utt, int beam, int max_edges, int H, int dk,
                             int window, float scale, void* stream) {
  PKA_REQUIRE(q && k_self && v_self && kcache && vcache && edge_prev && slot_edge && slot_active && out, PKA_EINVAL, "tree_attn: null pointer");
  PKA_REQUIRE(window >= 1 && window <= kTreeMaxWindow, PKA_EUNSUPPORTED, "tree_attn: window=%d (1..%d)", window, kTreeMaxWindow);
  PKA_REQUIRE(dk == 16 || dk == 32 || dk == 64 || dk == 128, PKA_EUNSUPPORTED, "tree_attn: head dim %d not in {16,32,64,128}", dk);
  const int n_slots = n_utt * beam;
  const int warps = n_slots * H;
  const int blocks = (warps + 3) / 4;
  cudaStream_t st = as_stream(stream);
#define TREE(DD) launch_k(tree_attn_kernel<DD>, blocks, 128, 0, st, q, k_self, v_self, ld, kcache, vcache, edge_prev, slot_edge, slot_active, out, n_slots, beam, max_edges, H, window, scale)
  if (dk == 16) TREE(16); else if (dk == 32) TREE(32); else if (dk == 64) TREE(64); else TREE(128);
#undef TREE
  return check_launch("tree_attn");
}

extern "C" int pka_beam_advance(const pka_beam_desc* d, const float* logits, int32_t* edge_prev, int32_t* edge_word,
                                int32_t* edge_depth, double* edge_weight, int32_t* n_edges, int32_t* beam_edges,
                                int32_t* beam_count, int32_t* slot_edge, int32_t* slot_active, int32_t* curr_length,
                                int32_t* done, int32_t* n_not_done, void* stream) {
  PKA_REQUIRE(d && logits && edge_prev && edge_word && edge_depth && edge_weight && n_edges && beam_edges && beam_count &&
              slot_edge && slot_active && curr_length && done && n_not_done, PKA_EINVAL, "beam_advance: null pointer");
  PKA_REQUIRE(d->n_utt > 0 && d->beam >= 1 && d->beam <= 64 && d->V >= 2, PKA_EUNSUPPORTED, "beam_advance: n_utt=%d beam=%d V=%d (beam<=64)", d->n_utt, d->beam, d->V);
  PKA_REQUIRE(d->max_edges >= 1 + d->beam * d->max_len, PKA_EINVAL, "beam_advance: max_edges=%d < 1+beam*max_len", d->max_edges);
  const size_t smem = sizeof(double) * ((size_t)d->beam * d->V + d->beam);
  PKA_REQUIRE(smem <= 200 * 1024, PKA_EUNSUPPORTED, "beam_advance: beam*V=%d candidates exceed shared memory", d->beam * d->V);
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(beam_advance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    PKA_REQUIRE(e == cudaSuccess, PKA_ELAUNCH, "beam_advance: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
  }
  launch_k(beam_advance_kernel, d->n_utt, 128, smem, as_stream(stream), *d, logits, edge_prev, edge_word, edge_depth, edge_weight, n_edges, beam_edges, beam_count, slot_edge, slot_active, curr_length, done, n_not_done);
  return check_launch("beam_advance");
}
