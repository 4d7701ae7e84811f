// Banded / key-padded multi-head attention, fp32-accumulating SIMT kernels (forward + backward), flash style:
// the [B*H, Lq, Lk] probability tensor of the reference (T/Modules.py:77-96) is never written to HBM, and the masks
// of T/Models.py:27-49 are evaluated as an in-kernel predicate
//        allowed(i,j) = key_mask[b,j] != 0  &&  (!band || i+start <= j <= i+end).
// Rows with no allowed key produce 0 output / 0 gradient (the reference's post-softmax masked_fill, T/Modules.py:90).
// Dropout on the probabilities: element (b,h,i,j) uses Philox index ((b*H+h)*Lq+i)*Lk8 + j with Lk8 = Lk rounded up to
// a multiple of 8, so that eight consecutive keys share one Philox call (the tensor-core kernels rely on this).
// Tiles entirely outside the band are skipped, so the decoder's (-10,0) band costs O(L*11) instead of O(L^2).
//
// Layout: q[B,Lq,ldq], k[B,Lk,ldk], v[B,Lk,ldv], out[B,Lq,ldo]; head h occupies columns [h*D,(h+1)*D) -- the
// reference's head-major batch replication (T/SubLayers.py:49-59) is replaced by a stride.
//
// Work split: one warp owns one "row item" (a query in fwd/dq, a key in dk/dv) and streams 32-wide tiles of the
// other axis through shared memory, lane = item inside the tile; reductions over the tile use warp shuffles.
#include "common.cuh"
#include <math_constants.h>

namespace pka {

constexpr int kAttWarps = 8;          // row items per CTA
constexpr int kAttTile = 32;          // streamed items per tile

struct AttnP {
  int B, H, Lq, Lk;
  int ldq, ldk, ldv, ldo;
  int use_band, start, end;
  float scale;
  pka_dropout drop;
};

__device__ __forceinline__ bool pair_allowed(const AttnP& p, const uint8_t* __restrict__ kmask_b, int i, int j) {
  if (i >= p.Lq || j >= p.Lk || j < 0) return false;
  if (kmask_b[j] == 0) return false;
  if (p.use_band && (j < i + p.start || j > i + p.end)) return false;
  return true;
}

// ------------------------------------------------------------------------------------------------ forward
template <typename T, int D>
__global__ void __launch_bounds__(kAttWarps * 32)
attn_fwd_kernel(const AttnP p, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                const uint8_t* __restrict__ kmask, T* __restrict__ out, float* __restrict__ lse) {
  pdl_wait();
  constexpr int R = D / 32 > 0 ? D / 32 : 1;       // output columns per lane
  __shared__ float Ks[kAttTile][D + 1];
  __shared__ float Vs[kAttTile][D];
  __shared__ float Qs[kAttWarps][D];
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * kAttWarps;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = i0 + warp;
  const uint8_t* kmask_b = kmask + (long long)b * p.Lk;
  const T* qb = q + (long long)b * p.Lq * p.ldq + h * D;
  const T* kb = k + (long long)b * p.Lk * p.ldk + h * D;
  const T* vb = v + (long long)b * p.Lk * p.ldv + h * D;

  for (int d = lane; d < D; d += 32) Qs[warp][d] = (i < p.Lq) ? to_f(qb[(long long)i * p.ldq + d]) : 0.f;

  int jlo = 0, jhi = p.Lk - 1;
  if (p.use_band) {
    jlo = max(0, i0 + p.start);
    jhi = min(p.Lk - 1, min(i0 + kAttWarps - 1, p.Lq - 1) + p.end);
  }
  DropCtx dc = make_drop(p.drop);
  float m_run = -CUDART_INF_F, l_run = 0.f, o[R];
#pragma unroll
  for (int r = 0; r < R; ++r) o[r] = 0.f;

  for (int j0 = (jlo / kAttTile) * kAttTile; j0 <= jhi; j0 += kAttTile) {
    __syncthreads();
    for (int e = threadIdx.x; e < kAttTile * D; e += kAttWarps * 32) {
      const int r = e / D, d = e % D, j = j0 + r;
      const bool in = j < p.Lk;
      Ks[r][d] = in ? to_f(kb[(long long)j * p.ldk + d]) : 0.f;
      Vs[r][d] = in ? to_f(vb[(long long)j * p.ldv + d]) : 0.f;
    }
    __syncthreads();
    const int j = j0 + lane;
    const bool ok = pair_allowed(p, kmask_b, i, j);
    float s = -CUDART_INF_F;
    if (ok) {
      float acc = 0.f;
#pragma unroll 16
      for (int d = 0; d < D; ++d) acc = fmaf(Qs[warp][d], Ks[lane][d], acc);
      s = acc * p.scale;
    }
    const float m_new = fmaxf(m_run, warp_max(s));
    if (m_new == -CUDART_INF_F) continue;            // warp-uniform: nothing allowed yet
    const float corr = (m_run == -CUDART_INF_F) ? 0.f : __expf(m_run - m_new);
    float pj = ok ? __expf(s - m_new) : 0.f;
    l_run = l_run * corr + warp_sum(pj);
    m_run = m_new;
    if (dc.p > 0.f && ok) {
      unsigned long long idx = (((unsigned long long)b * p.H + h) * p.Lq + i) * (unsigned long long)((p.Lk + 7) & ~7) + j;
      pj = dropout_keep(dc, idx) ? pj * dc.scale : 0.f;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) o[r] *= corr;
#pragma unroll 8
    for (int jj = 0; jj < kAttTile; ++jj) {
      const float w = __shfl_sync(0xffffffffu, pj, jj);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int d = lane + 32 * r;
        if (d < D) o[r] = fmaf(w, Vs[jj][d], o[r]);
      }
    }
  }
  if (i < p.Lq) {
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    T* ob = out + ((long long)b * p.Lq + i) * p.ldo + h * D;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int d = lane + 32 * r;
      if (d < D) ob[d] = from_f<T>(o[r] * inv);
    }
    if (lane == 0) lse[((long long)b * p.H + h) * p.Lq + i] = l_run > 0.f ? m_run + __logf(l_run) : -CUDART_INF_F;
  }
}

// ------------------------------------------------------------------------------------------------ backward: dq (+delta)
template <typename T, int D>
__global__ void __launch_bounds__(kAttWarps * 32)
attn_bwd_dq_kernel(const AttnP p, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                   const uint8_t* __restrict__ kmask, const T* __restrict__ out, const T* __restrict__ dout,
                   const float* __restrict__ lse, float* __restrict__ delta, T* __restrict__ dq) {
  pdl_wait();
  constexpr int R = D / 32 > 0 ? D / 32 : 1;
  __shared__ float Ks[kAttTile][D + 1];
  __shared__ float Vs[kAttTile][D + 1];
  __shared__ float Qs[kAttWarps][D];
  __shared__ float dOs[kAttWarps][D];
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * kAttWarps;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = i0 + warp;
  const uint8_t* kmask_b = kmask + (long long)b * p.Lk;
  const T* kb = k + (long long)b * p.Lk * p.ldk + h * D;
  const T* vb = v + (long long)b * p.Lk * p.ldv + h * D;

  float dl = 0.f;
  for (int d = lane; d < D; d += 32) {
    float qv = 0.f, dov = 0.f, ov = 0.f;
    if (i < p.Lq) {
      qv = to_f(q[((long long)b * p.Lq + i) * p.ldq + h * D + d]);
      dov = to_f(dout[((long long)b * p.Lq + i) * p.ldo + h * D + d]);
      ov = to_f(out[((long long)b * p.Lq + i) * p.ldo + h * D + d]);
    }
    Qs[warp][d] = qv; dOs[warp][d] = dov;
    dl = fmaf(dov, ov, dl);
  }
  dl = warp_sum(dl);                               // delta_i = dO_i . O_i
  const float lse_i = (i < p.Lq) ? lse[((long long)b * p.H + h) * p.Lq + i] : -CUDART_INF_F;
  if (i < p.Lq && lane == 0) delta[((long long)b * p.H + h) * p.Lq + i] = dl;

  int jlo = 0, jhi = p.Lk - 1;
  if (p.use_band) {
    jlo = max(0, i0 + p.start);
    jhi = min(p.Lk - 1, min(i0 + kAttWarps - 1, p.Lq - 1) + p.end);
  }
  DropCtx dc = make_drop(p.drop);
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;

  for (int j0 = (jlo / kAttTile) * kAttTile; j0 <= jhi; j0 += kAttTile) {
    __syncthreads();
    for (int e = threadIdx.x; e < kAttTile * D; e += kAttWarps * 32) {
      const int r = e / D, d = e % D, j = j0 + r;
      const bool in = j < p.Lk;
      Ks[r][d] = in ? to_f(kb[(long long)j * p.ldk + d]) : 0.f;
      Vs[r][d] = in ? to_f(vb[(long long)j * p.ldv + d]) : 0.f;
    }
    __syncthreads();
    const int j = j0 + lane;
    const bool ok = pair_allowed(p, kmask_b, i, j) && lse_i != -CUDART_INF_F;
    float ds = 0.f;
    if (ok) {
      float s = 0.f, dp = 0.f;
#pragma unroll 16
      for (int d = 0; d < D; ++d) {
        s = fmaf(Qs[warp][d], Ks[lane][d], s);
        dp = fmaf(dOs[warp][d], Vs[lane][d], dp);
      }
      const float pr = __expf(s * p.scale - lse_i);
      if (dc.p > 0.f) {
        unsigned long long idx = (((unsigned long long)b * p.H + h) * p.Lq + i) * (unsigned long long)((p.Lk + 7) & ~7) + j;
        dp = dropout_keep(dc, idx) ? dp * dc.scale : 0.f;
      }
      ds = pr * (dp - dl) * p.scale;
    }
#pragma unroll 8
    for (int jj = 0; jj < kAttTile; ++jj) {
      const float w = __shfl_sync(0xffffffffu, ds, jj);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int d = lane + 32 * r;
        if (d < D) acc[r] = fmaf(w, Ks[jj][d], acc[r]);
      }
    }
  }
  if (i < p.Lq) {
    T* dqb = dq + ((long long)b * p.Lq + i) * p.ldq + h * D;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int d = lane + 32 * r;
      if (d < D) dqb[d] = from_f<T>(acc[r]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward: dk, dv
template <typename T, int D>
__global__ void __launch_bounds__(kAttWarps * 32)
attn_bwd_dkv_kernel(const AttnP p, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                    const uint8_t* __restrict__ kmask, const T* __restrict__ dout, const float* __restrict__ lse,
                    const float* __restrict__ delta, T* __restrict__ dk, T* __restrict__ dv) {
  pdl_wait();
  constexpr int R = D / 32 > 0 ? D / 32 : 1;
  __shared__ float Qt[kAttTile][D + 1];
  __shared__ float dOt[kAttTile][D + 1];
  __shared__ float Kr[kAttWarps][D];
  __shared__ float Vr[kAttWarps][D];
  __shared__ float lse_t[kAttTile], del_t[kAttTile];
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * kAttWarps;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = j0 + warp;
  const uint8_t* kmask_b = kmask + (long long)b * p.Lk;
  const T* qb = q + (long long)b * p.Lq * p.ldq + h * D;
  const T* dob = dout + (long long)b * p.Lq * p.ldo + h * D;

  for (int d = lane; d < D; d += 32) {
    Kr[warp][d] = (j < p.Lk) ? to_f(k[((long long)b * p.Lk + j) * p.ldk + h * D + d]) : 0.f;
    Vr[warp][d] = (j < p.Lk) ? to_f(v[((long long)b * p.Lk + j) * p.ldv + h * D + d]) : 0.f;
  }
  int ilo = 0, ihi = p.Lq - 1;
  if (p.use_band) {                                // queries that can see keys j0 .. j0+W-1:  j-end <= i <= j-start
    ilo = max(0, j0 - p.end);
    ihi = min(p.Lq - 1, min(j0 + kAttWarps - 1, p.Lk - 1) - p.start);
  }
  DropCtx dc = make_drop(p.drop);
  float acck[R], accv[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { acck[r] = 0.f; accv[r] = 0.f; }

  for (int t0 = (ilo / kAttTile) * kAttTile; t0 <= ihi; t0 += kAttTile) {
    __syncthreads();
    for (int e = threadIdx.x; e < kAttTile * D; e += kAttWarps * 32) {
      const int r = e / D, d = e % D, i = t0 + r;
      const bool in = i < p.Lq;
      Qt[r][d] = in ? to_f(qb[(long long)i * p.ldq + d]) : 0.f;
      dOt[r][d] = in ? to_f(dob[(long long)i * p.ldo + d]) : 0.f;
    }
    if (threadIdx.x < kAttTile) {
      const int i = t0 + threadIdx.x;
      const bool in = i < p.Lq;
      lse_t[threadIdx.x] = in ? lse[((long long)b * p.H + h) * p.Lq + i] : -CUDART_INF_F;
      del_t[threadIdx.x] = in ? delta[((long long)b * p.H + h) * p.Lq + i] : 0.f;
    }
    __syncthreads();
    const int i = t0 + lane;
    const float lse_i = lse_t[lane];
    const bool ok = pair_allowed(p, kmask_b, i, j) && lse_i != -CUDART_INF_F;
    float pd = 0.f, ds = 0.f;
    if (ok) {
      float s = 0.f, dp = 0.f;
#pragma unroll 16
      for (int d = 0; d < D; ++d) {
        s = fmaf(Qt[lane][d], Kr[warp][d], s);
        dp = fmaf(dOt[lane][d], Vr[warp][d], dp);
      }
      const float pr = __expf(s * p.scale - lse_i);
      float mul = 1.f;
      if (dc.p > 0.f) {
        unsigned long long idx = (((unsigned long long)b * p.H + h) * p.Lq + i) * (unsigned long long)((p.Lk + 7) & ~7) + j;
        mul = dropout_keep(dc, idx) ? dc.scale : 0.f;
      }
      pd = pr * mul;                               // dropped probability: dV_j += pd * dO_i
      ds = pr * (dp * mul - del_t[lane]) * p.scale;
    }
#pragma unroll 8
    for (int ii = 0; ii < kAttTile; ++ii) {
      const float wv = __shfl_sync(0xffffffffu, pd, ii);
      const float wk = __shfl_sync(0xffffffffu, ds, ii);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int d = lane + 32 * r;
        if (d < D) {
          accv[r] = fmaf(wv, dOt[ii][d], accv[r]);
          acck[r] = fmaf(wk, Qt[ii][d], acck[r]);
        }
      }
    }
  }
  if (j < p.Lk) {
    T* dkb = dk + ((long long)b * p.Lk + j) * p.ldk + h * D;
    T* dvb = dv + ((long long)b * p.Lk + j) * p.ldv + h * D;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int d = lane + 32 * r;
      if (d < D) { dkb[d] = from_f<T>(acck[r]); dvb[d] = from_f<T>(accv[r]); }
    }
  }
}

// ------------------------------------------------------------------------------------------------ few queries, many keys
// Beam-search cross-attention (L/decode.py:85-87 re-runs it for every hypothesis; here the beam is the query axis): at
// most 16 queries per (utterance, head) against all T encoder frames.  The generic kernel above re-stages every K/V tile
// once per 8 queries; this one makes ONE pass over K and ONE over V per (utterance, head):
//   phase 1  lane = key: scores of all queries against its key row (Q broadcast from shared memory) -> S[q][j] in smem
//   phase 2  warp = query: max / sum over the keys, probabilities normalised in place
//   phase 3  lane = two output columns: out[q][:] += P[q][j] * V[j][:] with coalesced V rows, 8 warps interleave the keys
// fp32 throughout; rows without an allowed key give 0 / -inf like the generic kernel.
// Tried and removed (round 2): a key-split variant -- a cluster of 4 CTAs per (utterance, head), every live K / V row of
// a 128-key slice requested up front as a 256-byte bulk copy, partial results combined through distributed shared
// memory.  Correct, but SLOWER (46 vs 41 us at 125 utterances x 499 keys, 83 vs 62 us at 250): each of the 2000 short
// CTAs pays the mask -> expect_tx -> copy -> cluster-barrier -> remote-read latency chain for only 64 KB of traffic.
// The ceiling is not the copy bandwidth either: 12 fp32 queries per key are 6 FLOP per K/V byte, i.e. ~40 TFLOP/s of
// SIMT FMAs at HBM speed (more than half of the fp32 peak before a single shared-memory load is counted).
constexpr int kSqMaxQ = 16;
// NQ = queries rounded up to a multiple of 4 (beam 10 -> 12): the score / output loops are fully unrolled over NQ, so a
// 16-wide instantiation would spend a quarter of its FMAs on padding
template <int D, int NQ>
__global__ void __launch_bounds__(256, 2)
attn_fwd_smallq_kernel(const AttnP p, const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                       const uint8_t* __restrict__ kmask, float* __restrict__ out, float* __restrict__ lse, int LkP) {
  pdl_wait();
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;                                  // [16][D]
  float* S = Qs + NQ * D;                          // [NQ][LkP]
  float* red = S + NQ * LkP;                       // [8][NQ][D]
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint8_t* km = kmask + (long long)b * p.Lk;
  const float* kb = k + (long long)b * p.Lk * p.ldk + h * D;
  const float* vb = v + (long long)b * p.Lk * p.ldv + h * D;
  for (int e = threadIdx.x; e < NQ * D; e += 256) {
    const int qi = e / D, d = e - qi * D;
    Qs[e] = qi < p.Lq ? q[((long long)b * p.Lq + qi) * p.ldq + h * D + d] : 0.f;
  }
  __syncthreads();
  // ---- phase 1: scores
  for (int j = threadIdx.x; j < LkP; j += 256) {
    float s[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) s[qi] = 0.f;
    const bool ok = j < p.Lk && km[j] != 0;
    if (ok) {
      const float4* kr = reinterpret_cast<const float4*>(kb + (long long)j * p.ldk);
      float4 krow[D / 4];                          // the whole key row in flight before the first FMA
#pragma unroll
      for (int d4 = 0; d4 < D / 4; ++d4) krow[d4] = kr[d4];
#pragma unroll
      for (int d4 = 0; d4 < D / 4; ++d4) {
        const float4 kv = krow[d4];
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
          const float4 qv = *reinterpret_cast<const float4*>(Qs + qi * D + d4 * 4);
          s[qi] = fmaf(qv.x, kv.x, s[qi]); s[qi] = fmaf(qv.y, kv.y, s[qi]);
          s[qi] = fmaf(qv.z, kv.z, s[qi]); s[qi] = fmaf(qv.w, kv.w, s[qi]);
        }
      }
    }
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) S[qi * LkP + j] = ok ? s[qi] * p.scale : -CUDART_INF_F;
  }
  __syncthreads();
  // ---- phase 2: softmax over the keys, one warp per query
  for (int qi = warp; qi < NQ; qi += 8) {
    float* row = S + qi * LkP;
    float m = -CUDART_INF_F;
    for (int j = lane; j < LkP; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float l = 0.f;
    if (m != -CUDART_INF_F)
      for (int j = lane; j < LkP; j += 32) { const float e = __expf(row[j] - m); row[j] = e; l += e; }
    l = warp_sum(l);
    const float inv = l > 0.f ? 1.f / l : 0.f;
    for (int j = lane; j < LkP; j += 32) row[j] = (m != -CUDART_INF_F) ? row[j] * inv : 0.f;
    if (lane == 0 && qi < p.Lq) lse[((long long)b * p.H + h) * p.Lq + qi] = l > 0.f ? m + __logf(l) : -CUDART_INF_F;
  }
  __syncthreads();
  // ---- phase 3: out = P V
  constexpr int CPL = D / 32;                      // output columns per lane (D = 64 -> 2)
  float acc[NQ][CPL];
#pragma unroll
  for (int qi = 0; qi < NQ; ++qi)
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[qi][c] = 0.f;
  // eight V rows of this warp in flight per round (round 2: one row per round left a single load in flight, ~60
  // dependent round trips per CTA = 37 of the kernel's 40 us); the accumulation order over j is unchanged
  constexpr int kVU = 8;
  for (int j0 = warp; j0 < p.Lk; j0 += 8 * kVU) {
    float vv[kVU][CPL];
#pragma unroll
    for (int u = 0; u < kVU; ++u) {
      const int j = j0 + 8 * u;
#pragma unroll
      for (int c = 0; c < CPL; ++c) vv[u][c] = j < p.Lk ? vb[(long long)j * p.ldv + lane * CPL + c] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kVU; ++u) {
      const int j = j0 + 8 * u;
      if (j < p.Lk) {
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
          const float pj = S[qi * LkP + j];
#pragma unroll
          for (int c = 0; c < CPL; ++c) acc[qi][c] = fmaf(pj, vv[u][c], acc[qi][c]);
        }
      }
    }
  }
#pragma unroll
  for (int qi = 0; qi < NQ; ++qi)
#pragma unroll
    for (int c = 0; c < CPL; ++c) red[(warp * NQ + qi) * D + lane * CPL + c] = acc[qi][c];
  __syncthreads();
  for (int e = threadIdx.x; e < p.Lq * D; e += 256) {
    const int qi = e / D, d = e - qi * D;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[(w * NQ + qi) * D + d];
    out[((long long)b * p.Lq + qi) * p.ldo + h * D + d] = t;
  }
}

// ------------------------------------------------------------------------------------------------ optional probs dump
template <typename T>
__global__ void attn_probs_kernel(const AttnP p, int D, const T* __restrict__ q, const T* __restrict__ k,
                                  const uint8_t* __restrict__ kmask, const float* __restrict__ lse,
                                  float* __restrict__ probs) {
  pdl_wait();
  const long long total = (long long)p.B * p.H * p.Lq * p.Lk;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int j = e % p.Lk;
    const int i = (e / p.Lk) % p.Lq;
    const int h = (e / ((long long)p.Lk * p.Lq)) % p.H;
    const int b = e / ((long long)p.Lk * p.Lq * p.H);
    float pr = 0.f;
    const float l = lse[((long long)b * p.H + h) * p.Lq + i];
    if (pair_allowed(p, kmask + (long long)b * p.Lk, i, j) && l != -CUDART_INF_F) {
      const T* qr = q + ((long long)b * p.Lq + i) * p.ldq + h * D;
      const T* kr = k + ((long long)b * p.Lk + j) * p.ldk + h * D;
      float s = 0.f;
      for (int d = 0; d < D; ++d) s = fmaf(to_f(qr[d]), to_f(kr[d]), s);
      pr = __expf(s * p.scale - l);
    }
    probs[e] = pr;
  }
}

static int fill(AttnP& p, const pka_attn_desc* d, const char* who) {
  PKA_REQUIRE(d, PKA_EINVAL, "%s: null descriptor", who);
  PKA_REQUIRE(d->B > 0 && d->H > 0 && d->Lq > 0 && d->Lk > 0, PKA_EINVAL, "%s: bad sizes B=%d H=%d Lq=%d Lk=%d", who, d->B, d->H, d->Lq, d->Lk);
  PKA_REQUIRE(d->dk == d->dv, PKA_EUNSUPPORTED, "%s: d_k (%d) != d_v (%d) is not built", who, d->dk, d->dv);
  PKA_REQUIRE(d->dk == 16 || d->dk == 32 || d->dk == 64 || d->dk == 128, PKA_EUNSUPPORTED, "%s: head dim %d not in {16,32,64,128}", who, d->dk);
  PKA_REQUIRE(d->B <= 65535 && d->H <= 65535, PKA_EUNSUPPORTED, "%s: B or H exceeds grid limits", who);
  p.B = d->B; p.H = d->H; p.Lq = d->Lq; p.Lk = d->Lk;
  p.ldq = d->ldq; p.ldk = d->ldk; p.ldv = d->ldv; p.ldo = d->ldo;
  p.use_band = d->use_band; p.start = d->band_start; p.end = d->band_end;
  p.scale = d->scale; p.drop = d->drop;
  return PKA_OK;
}

#define PKA_ATT_DISPATCH(D_, CALL)                 \
  switch (D_) {                                    \
    case 16: { constexpr int DD = 16; CALL; } break;   \
    case 32: { constexpr int DD = 32; CALL; } break;   \
    case 64: { constexpr int DD = 64; CALL; } break;   \
    default: { constexpr int DD = 128; CALL; } break;  \
  }

// fp32, <= 16 queries, no band, no dropout, head dim 64, 16-byte aligned rows: the one-pass kernel above
static bool smallq_ok(const AttnP& p, int D, const void* q, const void* k, const void* v, int* LkP, int* smem) {
  if (D != 64 || p.Lq > kSqMaxQ || p.use_band || p.drop.p > 0.f) return false;
  if ((p.ldk & 3) || (p.ldv & 3) || !aligned16(k) || !aligned16(v) || !aligned16(q)) return false;
  *LkP = (p.Lk + 31) / 32 * 32;
  const int nq = (p.Lq + 3) / 4 * 4;
  *smem = (nq * D + nq * *LkP + 8 * nq * D) * (int)sizeof(float);
  return *smem <= 200 * 1024 && p.Lk >= 64;
}

template <typename T>
static int fwd_t(const AttnP& p, int D, const void* q, const void* k, const void* v, const uint8_t* km, void* out,
                 float* lse, float* probs, cudaStream_t st) {
  int LkP = 0, smem = 0;
  if (sizeof(T) == 4 && !probs && smallq_ok(p, D, q, k, v, &LkP, &smem)) {
    static bool smem_set = false;
    if (!smem_set) {
      cudaError_t e = cudaFuncSetAttribute(attn_fwd_smallq_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_smallq_kernel<64, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_smallq_kernel<64, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_smallq_kernel<64, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      PKA_REQUIRE(e == cudaSuccess, PKA_ELAUNCH, "attn_fwd: cannot opt in to shared memory: %s", cudaGetErrorString(e));
      smem_set = true;
    }
#define PKA_SMALLQ(NQ_) launch_k(attn_fwd_smallq_kernel<64, NQ_>, dim3(p.H, p.B), 256, (size_t)smem, st, p, (const float*)q, (const float*)k, (const float*)v, km, (float*)out, lse, LkP)
    if (p.Lq <= 4) PKA_SMALLQ(4); else if (p.Lq <= 8) PKA_SMALLQ(8); else if (p.Lq <= 12) PKA_SMALLQ(12); else PKA_SMALLQ(16);
#undef PKA_SMALLQ
    return check_launch("attn_fwd(smallq)");
  }
  dim3 grid((p.Lq + kAttWarps - 1) / kAttWarps, p.H, p.B), block(kAttWarps * 32);
  PKA_ATT_DISPATCH(D, (launch_k(attn_fwd_kernel<T, DD>, grid, block, 0, st, p, (const T*)q, (const T*)k, (const T*)v, km, (T*)out, lse)));
  int rc = check_launch("attn_fwd");
  if (rc) return rc;
  if (probs) {
    launch_k(attn_probs_kernel<T>, kNumSMs * 4, 256, 0, st, p, D, (const T*)q, (const T*)k, km, lse, probs);
    rc = check_launch("attn_probs");
  }
  return rc;
}

template <typename T>
static int bwd_t(const AttnP& p, int D, const void* q, const void* k, const void* v, const uint8_t* km,
                 const void* out, const void* dout, const float* lse, float* delta, void* dq, void* dk, void* dv,
                 cudaStream_t st) {
  dim3 block(kAttWarps * 32);
  dim3 gq((p.Lq + kAttWarps - 1) / kAttWarps, p.H, p.B);
  PKA_ATT_DISPATCH(D, (launch_k(attn_bwd_dq_kernel<T, DD>, gq, block, 0, st, p, (const T*)q, (const T*)k, (const T*)v, km, (const T*)out, (const T*)dout, lse, delta, (T*)dq)));
  int rc = check_launch("attn_bwd_dq");
  if (rc) return rc;
  dim3 gk((p.Lk + kAttWarps - 1) / kAttWarps, p.H, p.B);
  PKA_ATT_DISPATCH(D, (launch_k(attn_bwd_dkv_kernel<T, DD>, gk, block, 0, st, p, (const T*)q, (const T*)k, (const T*)v, km, (const T*)dout, lse, delta, (T*)dk, (T*)dv)));
  return check_launch("attn_bwd_dkv");
}

}  // namespace pka

extern "C" int pka_attn_fwd(const pka_attn_desc* d, int dtype, const void* q, const void* k, const void* v,
                            const uint8_t* key_mask, void* out, float* lse, float* probs_out, void* stream) {
  using namespace pka;
  AttnP p;
  int rc = fill(p, d, "attn_fwd");
  if (rc) return rc;
  PKA_REQUIRE(q && k && v && key_mask && out && lse, PKA_EINVAL, "attn_fwd: null pointer");
  if (dtype == PKA_F32) return fwd_t<float>(p, d->dk, q, k, v, key_mask, out, lse, probs_out, as_stream(stream));
  if (dtype == PKA_BF16) return fwd_t<__nv_bfloat16>(p, d->dk, q, k, v, key_mask, out, lse, probs_out, as_stream(stream));
  PKA_REQUIRE(false, PKA_EUNSUPPORTED, "attn_fwd: dtype %d", dtype);
}

extern "C" int pka_attn_bwd(const pka_attn_desc* d, int dtype, const void* q, const void* k, const void* v,
                            const uint8_t* key_mask, const void* out, const void* dout, const float* lse,
                            float* delta_ws, void* dq, void* dk, void* dv, void* stream) {
  using namespace pka;
  AttnP p;
  int rc = fill(p, d, "attn_bwd");
  if (rc) return rc;
  PKA_REQUIRE(q && k && v && key_mask && out && dout && lse && delta_ws && dq && dk && dv, PKA_EINVAL, "attn_bwd: null pointer");
  if (dtype == PKA_F32) return bwd_t<float>(p, d->dk, q, k, v, key_mask, out, dout, lse, delta_ws, dq, dk, dv, as_stream(stream));
  if (dtype == PKA_BF16) return bwd_t<__nv_bfloat16>(p, d->dk, q, k, v, key_mask, out, dout, lse, delta_ws, dq, dk, dv, as_stream(stream));
  PKA_REQUIRE(false, PKA_EUNSUPPORTED, "attn_bwd: dtype %d", dtype);
}
