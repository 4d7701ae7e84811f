// Fused front-end: [per-utterance CMVN] -> frame folding (sub-sampling by stacking) -> frame splicing, one pass over
// the padded feature batch, 128-bit coalesced accesses.  HBM-bound:
//   bytes = B*T*F*4 (read once: a CTA stages its frames, normalised, in shared memory and writes every spliced copy from
//   there) + B*(T/fold)*(n_ctx*F*fold)*sizeof(out) (write).
// Replaces: Kaldi apply-cmvn (P/run.sh:37-42, external binary) + fold_seq_and_mask (T/Models.py:51-65) +
// ConcatLayer (L/pytorch/TDNN.py:20-28).  Splicing pads with zeros beyond the *tensor* edge (row index outside
// [0, T/fold)), exactly like ConcatLayer's F.pad; padded frames inside the tensor are whatever the input holds (zeros),
// and stay zero under CMVN because the reference pads after feature extraction.
#include "common.cuh"

namespace pka {

// stats[b][0][f] = mean, stats[b][1][f] = 1/sqrt(var) (or 1 when only the mean is removed).
// One CTA per utterance.  The real frames of an utterance are one contiguous [n*F] run: a multiple of F/4 threads
// streams it as float4 (fully coalesced) and every thread keeps hitting the SAME four features (its stride is a
// multiple of F), so the per-feature sums live in registers (double) and meet once in shared memory, in thread order.
__global__ void __launch_bounds__(256)
cmvn_stats_kernel(const float* __restrict__ x, const int* __restrict__ lengths, float* __restrict__ stats, int T, int F,
                  int norm_vars) {
  pdl_wait();
  extern __shared__ double sm[];                  // [256][8]: 4 sums + 4 sums of squares per thread
  const int b = blockIdx.x, tid = threadIdx.x;
  int n = lengths[b];
  n = n < 0 ? 0 : (n > T ? T : n);
  const float* xb = x + (long long)b * T * F;
  const bool vec = (F % 4 == 0) && F / 4 <= 256 && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);
  const int per = vec ? F / 4 : F;                // distinct column groups
  const int nact = vec ? (256 / per) * per : (F <= 256 ? (256 / F) * F : 0);
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (nact > 0 && tid < nact) {
    if (vec) {
      const long long n4 = (long long)n * F / 4;
      auto acc = [&](const float4 v) {
        s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
        q[0] += (double)v.x * v.x; q[1] += (double)v.y * v.y; q[2] += (double)v.z * v.z; q[3] += (double)v.w * v.w;
      };
      const float4* x4 = reinterpret_cast<const float4*>(xb);
      long long e = tid;
      for (; e + 3LL * nact < n4; e += 4LL * nact) {           // four independent loads in flight per thread
        const float4 v0 = x4[e], v1 = x4[e + nact], v2 = x4[e + 2LL * nact], v3 = x4[e + 3LL * nact];
        acc(v0); acc(v1); acc(v2); acc(v3);
      }
      // (round 2, both measured at 2048 utterances x 450 frames x 40 and both slower than this loop's 38 us = 59 % of the
      //  copy bandwidth: a software-pipelined version with the next four loads requested before the accumulation -- 70
      //  registers, 3 CTAs per SM, 42 us; rounds of eight predicated loads with a speculative first round issued before
      //  the length is known -- 80 registers with spills, 53 us.  What is left is the serial part of every short CTA
      //  (length load, 4-5 dependent load rounds, the fixed-order double-precision finish) and the 3.5-wave tail.)
      for (; e < n4; e += nact) acc(x4[e]);
    } else {
      for (long long e = tid; e < (long long)n * F; e += nact) { const double v = xb[e]; s[0] += v; q[0] += v * v; }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) { sm[tid * 8 + i] = s[i]; sm[tid * 8 + 4 + i] = q[i]; }
  __syncthreads();
  for (int f = tid; f < F; f += 256) {
    double ss = 0.0, qq = 0.0;
    if (nact > 0) {
      const int grp = vec ? f / 4 : f, sub = vec ? f % 4 : 0;
      for (int t = grp; t < nact; t += per) { ss += sm[t * 8 + sub]; qq += sm[t * 8 + 4 + sub]; }
    } else {                                      // very wide features: plain strided pass
      for (int t = 0; t < n; ++t) { const double v = xb[(long long)t * F + f]; ss += v; qq += v * v; }
    }
    double mean = n > 0 ? ss / n : 0.0, istd = 1.0;
    if (norm_vars && n > 0) {
      double var = qq / n - mean * mean;
      if (var < 1e-20) var = 1e-20;
      istd = 1.0 / sqrt(var);
    }
    stats[((long long)b * 2 + 0) * F + f] = (float)mean;
    stats[((long long)b * 2 + 1) * F + f] = (float)istd;
  }
}

struct FrontP {
  int B, T, F, fold, n_ctx, cmvn;
  int ctx[PKA_MAX_CTX];
};

// One warp per output row (b, t): lane j owns output columns [VEC*j, VEC*j + VEC) (+ 32*VEC strides for wide rows), so
// the column -> (context, frame-in-fold, feature) split is a handful of 32-bit operations and every access is one
// 16-byte (fp32 in / bf16 out with VEC = 8) or 16-byte / 16-byte (VEC = 4, fp32 out) coalesced vector.
// One thread per (output row, VEC-wide column group), all 32 lanes of every warp busy, two items in flight per thread.
// The column-group -> (context shift, frame inside the fold, feature) split does not depend on the row: it is tabulated
// once per CTA in shared memory, so an item costs two 32-bit divisions, one 16/32-byte load and one 16-byte store.
constexpr int kFeMaxGroups = 1024;
template <typename To, int VEC>
__global__ void __launch_bounds__(256)
frontend_kernel(const FrontP p, const float* __restrict__ x, const int* __restrict__ lengths,
                const float* __restrict__ stats, To* __restrict__ out) {
  pdl_wait();
  __shared__ short g_shift[kFeMaxGroups], g_fr[kFeMaxGroups], g_f[kFeMaxGroups];
  const int Tf = p.T / p.fold, Ff = p.F * p.fold, W = p.n_ctx * Ff;
  const unsigned G = (unsigned)(W / VEC);
  for (unsigned g = threadIdx.x; g < G; g += blockDim.x) {
    const int col = (int)g * VEC, c = col / Ff, fp = col - c * Ff, fr = fp / p.F;
    g_shift[g] = (short)p.ctx[c]; g_fr[g] = (short)fr; g_f[g] = (short)(fp - fr * p.F);   // VEC consecutive f share a frame
  }
  __syncthreads();
  const unsigned long long total = (unsigned long long)p.B * Tf * G;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  auto item = [&](unsigned long long idx) {
    const unsigned row = (unsigned)(idx / G), g = (unsigned)(idx - (unsigned long long)row * G);
    const unsigned b = row / (unsigned)Tf;
    const int t = (int)(row - b * (unsigned)Tf);
    const int ts = t + g_shift[g], f = g_f[g];
    float v[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = 0.f;
    if (ts >= 0 && ts < Tf) {
      const int frame = ts * p.fold + g_fr[g];
      const float* src = x + ((long long)b * p.T + frame) * p.F + f;
      if (VEC == 1) v[0] = src[0];
      else {
#pragma unroll
        for (int i = 0; i < VEC; i += 4) {
          const float4 q = *reinterpret_cast<const float4*>(src + i);
          v[i] = q.x; v[(i + 1) % VEC] = q.y; v[(i + 2) % VEC] = q.z; v[(i + 3) % VEC] = q.w;
        }
      }
      if (p.cmvn) {
        const bool real = frame < lengths[b];
        const float* mu = stats + ((long long)b * 2 + 0) * p.F + f;
        const float* is = stats + ((long long)b * 2 + 1) * p.F + f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = real ? (v[i] - mu[i]) * is[i] : 0.f;
      }
    }
    To* dst = out + (long long)row * W + g * VEC;
    if (VEC == 1) dst[0] = from_f<To>(v[0]);
    else if (VEC == 8 && sizeof(To) == 2) {        // one 16-byte store of 8 bf16
      uint4 pk;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1 % VEC]), h1 = __floats2bfloat162_rn(v[2 % VEC], v[3 % VEC]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4 % VEC], v[5 % VEC]), h3 = __floats2bfloat162_rn(v[6 % VEC], v[7 % VEC]);
      pk.x = *(uint32_t*)&h0; pk.y = *(uint32_t*)&h1; pk.z = *(uint32_t*)&h2; pk.w = *(uint32_t*)&h3;
      *reinterpret_cast<uint4*>(dst) = pk;
    } else {
#pragma unroll
      for (int i = 0; i < VEC; i += 4) st4(dst + i, make_float4(v[i], v[(i + 1) % VEC], v[(i + 2) % VEC], v[(i + 3) % VEC]));
    }
  };
  unsigned long long idx = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  for (; idx + stride < total; idx += 2 * stride) { item(idx); item(idx + stride); }
  if (idx < total) item(idx);
}

// Tile kernel (the one that runs; frontend_kernel above remains for layouts it cannot take: F % 4 != 0, tiles beyond the
// 48 KB of static-limit shared memory).  A CTA owns TT consecutive output frames of one utterance:
//   1. stage: the (TT + ctx span) * fold input frames it needs are ONE contiguous run of the utterance -> coalesced float4
//      loads, CMVN applied once per input element (mean / inverse deviation from shared memory), zeros outside [0, T);
//   2. emit: a warp per output frame, a lane per VEC-wide column group; the column -> (context, frame in the fold,
//      feature) split is done once per lane, an item is two LDS.128 + one 16-byte store, no division, no global re-read.
//      The tile is stored in 16-byte units with unit u at u ^ ((u >> 3) & 1): a lane that reads units 2g and 2g+1 (its 8
//      floats) would otherwise make every LDS.128 of a warp hit only the even (or odd) units of each 128-byte bank row
//      (2-way conflicts, L1/TEX was the busiest unit at 71 %); with the swap in every other bank row the 8 lanes of a
//      quarter warp touch 8 different units.
// Before (round 2, ncu, 2048 x 499 x 40 -> 200 bf16 columns): 111 warp instructions per item (64-bit divisions) and five
// L1 reads of every input element, the CMVN arithmetic repeated per copy: 66 % of the copy bandwidth without, 41 % with CMVN.
template <typename To, int VEC>
__global__ void __launch_bounds__(256)
frontend_tile_kernel(const FrontP p, int TT, int tiles, int smin, int smax, const float* __restrict__ x,
                     const int* __restrict__ lengths, const float* __restrict__ stats, To* __restrict__ out) {
  pdl_wait();
  extern __shared__ __align__(128) float fe_sm[];
  const int F = p.F, fold = p.fold, Tf = p.T / fold, Ff = F * fold, W = p.n_ctx * Ff, G = W / VEC;
  const int b = blockIdx.x / tiles, t0 = (blockIdx.x - b * tiles) * TT;
  const int t1 = t0 + TT < Tf ? t0 + TT : Tf;
  const int f_lo = (t0 + smin) * fold;             // first staged input frame (may be negative: zero fill)
  const int n_fr = (TT + smax - smin) * fold;
  float* st = fe_sm;                               // [mean | 1/sigma][F]   (CMVN only)
  float* tile = fe_sm + (p.cmvn ? (2 * F + 31) / 32 * 32 : 0);      // [n_fr][F] in swizzled 16-byte units, 128-byte aligned
  const int tid = threadIdx.x;
  int len = p.T;
  if (p.cmvn) {
    for (int i = tid; i < 2 * F; i += 256) st[i] = stats[(long long)b * 2 * F + i];
    len = lengths[b];
    __syncthreads();
  }
  {
    const int F4 = F >> 2, total4 = n_fr * F4;
    const float* xb = x + (long long)b * p.T * F;
    int fr = tid / F4, f4 = tid - fr * F4;         // (frame, vector) of element tid; advanced without divisions
    const int dfr = 256 / F4, df4 = 256 - dfr * F4;
    for (int j = tid; j < total4; j += 256) {
      const int frame = f_lo + fr;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (frame >= 0 && frame < p.T) {
        v = *reinterpret_cast<const float4*>(xb + (long long)frame * F + f4 * 4);
        if (p.cmvn) {
          if (frame < len) {
            const float4 mu = *reinterpret_cast<const float4*>(st + f4 * 4);
            const float4 is = *reinterpret_cast<const float4*>(st + F + f4 * 4);
            v.x = (v.x - mu.x) * is.x; v.y = (v.y - mu.y) * is.y; v.z = (v.z - mu.z) * is.z; v.w = (v.w - mu.w) * is.w;
          } else {
            v = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      const int u = fr * F4 + f4;
      reinterpret_cast<float4*>(tile)[u ^ ((u >> 3) & 1)] = v;
      fr += dfr; f4 += df4;
      if (f4 >= F4) { f4 -= F4; ++fr; }
    }
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int g = lane; g < G; g += 32) {
    const int col = g * VEC, c = col / Ff, fp = col - c * Ff, frr = fp / F, f = fp - frr * F;
    const int shift = p.ctx[c];
    const int u0 = ((frr - f_lo) * F + f) >> 2;   // first 16-byte unit of this lane's columns in tile row ts = 0
    const float4* tile4 = reinterpret_cast<const float4*>(tile);
#pragma unroll 4
    for (int t = t0 + warp; t < t1; t += 8) {
      const int ts = t + shift;
      float v[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) v[i] = 0.f;
      if (ts >= 0 && ts < Tf) {
        const int u = u0 + ts * fold * (F >> 2);
#pragma unroll
        for (int i = 0; i < VEC; i += 4) {
          const int ui = u + (i >> 2);
          const float4 q = tile4[ui ^ ((ui >> 3) & 1)];
          v[i] = q.x; v[i + 1] = q.y; v[i + 2] = q.z; v[i + 3] = q.w;
        }
      }
      To* dst = out + ((long long)b * Tf + t) * W + col;
      if (VEC == 8) {                              // bf16: one 16-byte store of 8
        uint4 pk;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4 % VEC], v[5 % VEC]), h3 = __floats2bfloat162_rn(v[6 % VEC], v[7 % VEC]);
        pk.x = *(uint32_t*)&h0; pk.y = *(uint32_t*)&h1; pk.z = *(uint32_t*)&h2; pk.w = *(uint32_t*)&h3;
        *reinterpret_cast<uint4*>(dst) = pk;
      } else {
        st4(dst, make_float4(v[0], v[1], v[2], v[3]));
      }
    }
  }
}

}  // namespace pka

extern "C" int pka_frontend_fwd(const float* feats, const int32_t* lengths, void* out, int out_dtype, int B, int T, int F,
                                int fold, const int32_t* ctx_host, int n_ctx, int cmvn_mode, float* stats_ws,
                                void* stream) {
  using namespace pka;
  PKA_REQUIRE(feats && out && ctx_host, PKA_EINVAL, "frontend_fwd: null pointer");
  PKA_REQUIRE(B > 0 && T > 0 && F > 0 && fold >= 1 && T / fold >= 1, PKA_EINVAL, "frontend_fwd: B=%d T=%d F=%d fold=%d", B, T, F, fold);
  PKA_REQUIRE(n_ctx >= 1 && n_ctx <= PKA_MAX_CTX, PKA_EUNSUPPORTED, "frontend_fwd: n_ctx=%d (max %d)", n_ctx, PKA_MAX_CTX);
  PKA_REQUIRE(cmvn_mode >= 0 && cmvn_mode <= 2, PKA_EINVAL, "frontend_fwd: cmvn_mode=%d", cmvn_mode);
  PKA_REQUIRE(cmvn_mode == 0 || (lengths && stats_ws), PKA_EINVAL, "frontend_fwd: CMVN needs lengths and stats_ws");
  cudaStream_t st = as_stream(stream);
  if (cmvn_mode) {
    launch_k(cmvn_stats_kernel, B, 256, 256 * 8 * sizeof(double), st, feats, lengths, stats_ws, T, F, cmvn_mode == 2);
    int rc = check_launch("cmvn_stats");
    if (rc) return rc;
  }
  FrontP p;
  p.B = B; p.T = T; p.F = F; p.fold = fold; p.n_ctx = n_ctx; p.cmvn = cmvn_mode;
  for (int i = 0; i < PKA_MAX_CTX; ++i) p.ctx[i] = i < n_ctx ? ctx_host[i] : 0;
  const long long W = (long long)n_ctx * F * fold;
  PKA_REQUIRE(W <= kFeMaxGroups, PKA_EUNSUPPORTED, "frontend_fwd: spliced row width %lld exceeds %d", W, kFeMaxGroups);
  const long long items = (long long)B * (T / fold) * (W / 4);
  long long blocks = (items + 511) / 512;
  if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
  if (blocks < 1) blocks = 1;
  const bool v4 = (F % 4 == 0) && aligned16(feats) && aligned16(out);
  const bool v8 = v4 && (F % 8 == 0);
  if (v4) {                                        // shared-memory tile kernel whenever the layout allows it
    int smin = ctx_host[0], smax = ctx_host[0];
    for (int i = 1; i < n_ctx; ++i) { smin = ctx_host[i] < smin ? ctx_host[i] : smin; smax = ctx_host[i] > smax ? ctx_host[i] : smax; }
    const int Tf = T / fold;
    int TT = 64;
    // (the unit swap may touch the unit after the last one: round the tile up to a whole 128-byte bank row)
    auto smem_of = [&](int tt) { return (((long long)(tt + smax - smin) * fold * F + 31) / 32 * 32 + (cmvn_mode ? (2 * F + 31) / 32 * 32 : 0)) * 4; };
    while (TT > 8 && smem_of(TT) > 48 * 1024) TT >>= 1;
    const long long tiles = (Tf + TT - 1) / TT;
    if (smem_of(TT) <= 48 * 1024 && (long long)B * tiles < (1LL << 31)) {
      const unsigned grid = (unsigned)((long long)B * tiles);
      const size_t smem = (size_t)smem_of(TT);
#define PKA_FT(To, V) launch_k(frontend_tile_kernel<To, V>, grid, 256, smem, st, p, TT, (int)tiles, smin, smax, feats, lengths, stats_ws, (To*)out)
      if (out_dtype == PKA_F32) PKA_FT(float, 4);
      else if (out_dtype == PKA_BF16) { if (v8) PKA_FT(__nv_bfloat16, 8); else PKA_FT(__nv_bfloat16, 4); }
      else PKA_REQUIRE(false, PKA_EUNSUPPORTED, "frontend_fwd: out dtype %d", out_dtype);
#undef PKA_FT
      return check_launch("frontend (tile)");
    }
  }
#define PKA_FE(To, V) launch_k(frontend_kernel<To, V>, (int)blocks, 256, 0, st, p, feats, lengths, stats_ws, (To*)out)
  if (out_dtype == PKA_F32) {
    if (v4) PKA_FE(float, 4); else PKA_FE(float, 1);
  } else if (out_dtype == PKA_BF16) {
    if (v8) PKA_FE(__nv_bfloat16, 8); else if (v4) PKA_FE(__nv_bfloat16, 4); else PKA_FE(__nv_bfloat16, 1);
  } else PKA_REQUIRE(false, PKA_EUNSUPPORTED, "frontend_fwd: out dtype %d", out_dtype);
#undef PKA_FE
  return check_launch("frontend");
}
