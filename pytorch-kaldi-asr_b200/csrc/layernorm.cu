// Fused [dropout] + residual-add + LayerNormalization (forward / backward), one warp per row, warp-shuffle reductions.
// The normalisation is the reference's own (T/Modules.py:42-51), NOT F.layer_norm:
//      y = (z - mean(z)) / (std_unbiased(z) + eps) * a + b,   z = dropout(x) + residual,  eps = 1e-3 added to sigma.
// HBM-bound: forward moves 3*rows*D elements (x, residual in; y out), backward 5 (dy, x, residual in; dres[, dx] out).
#include "common.cuh"

namespace pka {

constexpr int kLnWarps = 8;

// ---- EV consecutive elements <-> one raw 16-byte (fp32 x4, bf16 x8) or 8-byte (bf16 x4) vector <-> fp32 registers.
// A loaded vector stays RAW (packed) while it is in flight: the loads of the next row iteration are issued before the
// arithmetic of the current one, and a packed bf16 vector costs half the registers of its unpacked values.
template <typename T, int EV> struct RawOf;
template <> struct RawOf<float, 4> { using type = float4; };
template <> struct RawOf<__nv_bfloat16, 8> { using type = uint4; };
template <> struct RawOf<__nv_bfloat16, 4> { using type = uint2; };

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t bf_pack(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void unpackv(const float4& r, float (&v)[4]) { v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
__device__ __forceinline__ void unpackv(const uint4& r, float (&v)[8]) {
  v[0] = bf_lo(r.x); v[1] = bf_hi(r.x); v[2] = bf_lo(r.y); v[3] = bf_hi(r.y);
  v[4] = bf_lo(r.z); v[5] = bf_hi(r.z); v[6] = bf_lo(r.w); v[7] = bf_hi(r.w);
}
__device__ __forceinline__ void unpackv(const uint2& r, float (&v)[4]) {
  v[0] = bf_lo(r.x); v[1] = bf_hi(r.x); v[2] = bf_lo(r.y); v[3] = bf_hi(r.y);
}
__device__ __forceinline__ void packv(const float (&v)[4], float4& r) { r = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void packv(const float (&v)[8], uint4& r) {
  r.x = bf_pack(v[0], v[1]); r.y = bf_pack(v[2], v[3]); r.z = bf_pack(v[4], v[5]); r.w = bf_pack(v[6], v[7]);
}
__device__ __forceinline__ void packv(const float (&v)[4], uint2& r) { r.x = bf_pack(v[0], v[1]); r.y = bf_pack(v[2], v[3]); }
// the value a stored element reads back as (bf16 rounding; identity for fp32)
template <typename T> __device__ __forceinline__ float stored(float v) { return to_f(from_f<T>(v)); }

// keep bits of EV consecutive elements starting at element index e0 (e0 % EV == 0): bit i = element e0 + i
template <int EV> __device__ __forceinline__ uint32_t drop_bitsv(const DropCtx& dc, unsigned long long e0) {
  const uint32_t b = dropout_bits8(dc, e0 >> 3);
  return EV == 8 ? b : (b >> (uint32_t)(((e0 >> 2) & 1ull) * 4ull)) & 0xfu;
}
// sum over the LPR lanes that share a row (LPR = power of two; xor offsets below LPR stay inside the lane group)
template <int LPR> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int N> __device__ __forceinline__ void ld_f32v(const float* p, float (&v)[N]) {
#pragma unroll
  for (int i = 0; i < N; i += 4) {
    const float4 q = *reinterpret_cast<const float4*>(p + i);
    v[i] = q.x; v[i + 1] = q.y; v[i + 2] = q.z; v[i + 3] = q.w;
  }
}

// Geometry of both kernels: a row is shared by LPR lanes (32, 16 or 8: a 128-wide bf16 row is 16 lanes x 16 bytes, so a
// warp works on 32/LPR rows at once and the shuffle reductions are log2(LPR) deep), each lane holds VPL vectors of EV
// elements (row capacity LPR*VPL*EV >= D).  The grid is persistent: a warp walks its rows with a grid stride, the gain /
// offset vectors live in registers for the whole walk (they cost 2 KB of L1 traffic per 1.5 KB of row data when they are
// re-read per row), and the raw vectors of the NEXT iteration are requested before the arithmetic of the current one.
// Measured before this layout (round 2, ncu, 223 552 x 256): bf16 forward 237 warp instructions per row at 45 % achieved
// occupancy (one-shot CTAs) = 56 % of the copy bandwidth; bf16 backward 99 registers -> 16 resident warps per SM with one
// row in flight each = 41 %.
template <typename T, int EV, int LPR, int VPL, bool DROP>
__global__ void __launch_bounds__(kLnWarps * 32, (VPL * EV <= 8) ? 4 : ((VPL * EV <= 16) ? 2 : 1))
add_ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ a,
                  const float* __restrict__ bta, T* __restrict__ y, float* __restrict__ mean_o,
                  float* __restrict__ rinv_o, int rows, int D, float eps, const pka_dropout drop) {
  using Raw = typename RawOf<T, EV>::type;
  constexpr int RPI = 32 / LPR;                       // rows per warp iteration
  constexpr bool HOIST = VPL * EV <= 16;              // gain / offset in registers
  constexpr bool PREFETCH = VPL * EV * sizeof(T) <= 64;
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  DropCtx dc;
  if (DROP) dc = make_drop(drop);
  const bool has_res = res != nullptr;
  const float inv_D = 1.f / (float)D, inv_Dm1 = 1.f / (float)(D - 1);   // (a full-precision division per row otherwise)
  float av[HOIST ? VPL : 1][EV], bv[HOIST ? VPL : 1][EV];
  if (HOIST) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = (v * LPR + l) * EV;
#pragma unroll
      for (int i = 0; i < EV; ++i) { av[v][i] = 0.f; bv[v][i] = 0.f; }
      if (c < D) { ld_f32v<EV>(a + c, av[v]); ld_f32v<EV>(bta + c, bv[v]); }
    }
  }
  const long long stride = (long long)gridDim.x * (kLnWarps * RPI);
  const long long base0 = ((long long)blockIdx.x * kLnWarps + warp) * RPI;
  Raw xr[VPL], rr[VPL];
  auto fetch = [&](long long r) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = (v * LPR + l) * EV;
      xr[v] = Raw{}; rr[v] = Raw{};
      if (r < rows && c < D) {
        xr[v] = *reinterpret_cast<const Raw*>(x + r * D + c);
        if (has_res) rr[v] = *reinterpret_cast<const Raw*>(res + r * D + c);
      }
    }
  };
  if (PREFETCH) fetch(base0 + sub);
  for (long long base = base0; base < rows; base += stride) {       // warp-uniform trip count
    const long long row = base + sub;
    const bool ok = row < rows;
    if (!PREFETCH) fetch(row);
    float z[VPL][EV];
    float sum = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = (v * LPR + l) * EV;
      float t[EV];
      unpackv(xr[v], z[v]);
      unpackv(rr[v], t);
      if (DROP) {
        if (ok && c < D) {
          const uint32_t kb = drop_bitsv<EV>(dc, (unsigned long long)row * D + c);
#pragma unroll
          for (int i = 0; i < EV; ++i) z[v][i] *= ((kb >> i) & 1u) ? dc.scale : 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < EV; ++i) { z[v][i] += t[i]; sum += z[v][i]; }   // columns >= D hold zeros
    }
    if (PREFETCH) fetch(row + stride);                               // in flight during the arithmetic below
    const float mean = group_sum<LPR>(sum) * inv_D;
    float sq = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = (v * LPR + l) * EV;
      if (c < D) {
#pragma unroll
        for (int i = 0; i < EV; ++i) { z[v][i] -= mean; sq = fmaf(z[v][i], z[v][i], sq); }
      }
    }
    // MUFU.RSQ / MUFU.RCP without the IEEE fix-up sequences (2 ulp / 1 ulp: far inside the stated tolerances; the exact
    // sqrt + division cost 25 of the ~165 instructions of a row, and this kernel is issue-bound in bf16)
    const float var = group_sum<LPR>(sq) * inv_Dm1;
    const float sigma = var > 1e-30f ? var * rsqrtf(var) : 0.f;
    const float rinv = __fdividef(1.f, sigma + eps);
    if (ok) {
      T* yr = y + row * D;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int c = (v * LPR + l) * EV;
        if (c < D) {
          float o[EV];
          if (HOIST) {
#pragma unroll
            for (int i = 0; i < EV; ++i) o[i] = z[v][i] * rinv * av[v][i] + bv[v][i];
          } else {
            float aa[EV], bb[EV];
            ld_f32v<EV>(a + c, aa); ld_f32v<EV>(bta + c, bb);
#pragma unroll
            for (int i = 0; i < EV; ++i) o[i] = z[v][i] * rinv * aa[i] + bb[i];
          }
          Raw pk;
          packv(o, pk);
          *reinterpret_cast<Raw*>(yr + c) = pk;
        }
      }
      if (l == 0) { mean_o[row] = mean; rinv_o[row] = rinv; }
    }
  }
}

// backward: dz_i = rinv*(g_i - mean(g)) - c_i * rinv^2 * sum(g*c) / ((D-1)*sigma),  g = dy*a, c = z-mean
//           dres = dz,  dx = keep*dz/(1-p),  da = sum_rows dy*c*rinv,  db = sum_rows dy
template <typename T, int EV, int LPR, int VPL, bool DROP>
__global__ void __launch_bounds__(kLnWarps * 32, (VPL * EV <= 8) ? ((DROP || sizeof(T) == 4) ? 2 : 3) : ((VPL * EV * sizeof(T) <= 32) ? 2 : 1))
add_ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ res,
                  const float* __restrict__ a, const float* __restrict__ mean_i, const float* __restrict__ rinv_i,
                  T* __restrict__ dx, T* __restrict__ dres, float* __restrict__ dab_ws, int rows, int D, float eps,
                  const pka_dropout drop) {
  using Raw = typename RawOf<T, EV>::type;
  constexpr int RPI = 32 / LPR;
  constexpr int CW = LPR * VPL * EV;                  // column capacity of a lane group
  constexpr bool HOIST = VPL * EV <= 8;               // gain in registers (wider rows re-read it from L1: the three
                                                      // accumulator sets already take 3 registers per element)
  constexpr bool PREFETCH = VPL * EV * sizeof(T) <= 32;
  pdl_wait();
  __shared__ float red[kLnWarps * RPI][3][CW];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  DropCtx dc;
  if (DROP) dc = make_drop(drop);
  const bool has_res = res != nullptr;
  const float inv_D = 1.f / (float)D, inv_Dm1 = 1.f / (float)(D - 1);
  float da_acc[VPL][EV], db_acc[VPL][EV], dxs_acc[VPL][EV];       // dxs: column sums of the dx this kernel writes
  float av[HOIST ? VPL : 1][EV];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c = (v * LPR + l) * EV;
#pragma unroll
    for (int i = 0; i < EV; ++i) { da_acc[v][i] = 0.f; db_acc[v][i] = 0.f; dxs_acc[v][i] = 0.f; }
    if (HOIST) {
#pragma unroll
      for (int i = 0; i < EV; ++i) av[v][i] = 0.f;
      if (c < D) ld_f32v<EV>(a + c, av[v]);
    }
  }
  const long long stride = (long long)gridDim.x * (kLnWarps * RPI);
  const long long base0 = ((long long)blockIdx.x * kLnWarps + warp) * RPI;
  Raw gr[VPL], xr[VPL], rr[VPL];
  float mean_n = 0.f, rinv_n = 1.f;
  auto fetch = [&](long long r) {
    const bool in = r < rows;
    mean_n = in ? mean_i[r] : 0.f;
    rinv_n = in ? rinv_i[r] : 1.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = (v * LPR + l) * EV;
      gr[v] = Raw{}; xr[v] = Raw{}; rr[v] = Raw{};
      if (in && c < D) {
        gr[v] = *reinterpret_cast<const Raw*>(dy + r * D + c);
        xr[v] = *reinterpret_cast<const Raw*>(x + r * D + c);
        if (has_res) rr[v] = *reinterpret_cast<const Raw*>(res + r * D + c);
      }
    }
  };
  if (PREFETCH) fetch(base0 + sub);
  for (long long base = base0; base < rows; base += stride) {       // warp-uniform trip count
    const long long row = base + sub;
    const bool ok = row < rows;
    if (!PREFETCH) fetch(row);
    const float mean = mean_n, rinv = rinv_n;
    const float sigma = 1.f / rinv - eps;
    float c_[VPL][EV], g[VPL][EV];
    uint32_t kb[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int col = (v * LPR + l) * EV;
      float dyv[EV], rv[EV];
      unpackv(gr[v], dyv);
      unpackv(xr[v], c_[v]);
      unpackv(rr[v], rv);
      kb[v] = 0xffu;
      const bool live = ok && col < D;
      if (DROP) {
        if (live) kb[v] = drop_bitsv<EV>(dc, (unsigned long long)row * D + col);
#pragma unroll
        for (int i = 0; i < EV; ++i) c_[v][i] *= ((kb[v] >> i) & 1u) ? dc.scale : 0.f;
      }
      float aa[EV];
      if (HOIST) {
#pragma unroll
        for (int i = 0; i < EV; ++i) aa[i] = av[v][i];
      } else {
#pragma unroll
        for (int i = 0; i < EV; ++i) aa[i] = 0.f;
        if (col < D) ld_f32v<EV>(a + col, aa);
      }
#pragma unroll
      for (int i = 0; i < EV; ++i) {
        c_[v][i] = live ? c_[v][i] + rv[i] - mean : 0.f;            // dead lanes / rows contribute exact zeros
        g[v][i] = dyv[i] * aa[i];
        s1 += g[v][i];
        s2 = fmaf(g[v][i], c_[v][i], s2);
        da_acc[v][i] = fmaf(dyv[i] * c_[v][i], rinv, da_acc[v][i]);
        db_acc[v][i] += dyv[i];
      }
    }
    if (PREFETCH) fetch(row + stride);                               // in flight during the reductions and the stores
    s1 = group_sum<LPR>(s1); s2 = group_sum<LPR>(s2);
    const float gm = s1 * inv_D;
    const float kf = sigma > 0.f ? rinv * rinv * s2 * inv_Dm1 / sigma : 0.f;
    if (ok) {
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int col = (v * LPR + l) * EV;
        if (col < D) {
          float dz[EV];
#pragma unroll
          for (int i = 0; i < EV; ++i) dz[i] = rinv * (g[v][i] - gm) - c_[v][i] * kf;
          Raw pk;
          if (dres) { packv(dz, pk); *reinterpret_cast<Raw*>(dres + row * D + col) = pk; }
          if (dx) {
            if (DROP) {
#pragma unroll
              for (int i = 0; i < EV; ++i) dz[i] *= ((kb[v] >> i) & 1u) ? dc.scale : 0.f;
            }
            packv(dz, pk);
            *reinterpret_cast<Raw*>(dx + row * D + col) = pk;
          }
          // bias gradient of the linear layer below = column sums of exactly what it will read (the stored, rounded dx)
#pragma unroll
          for (int i = 0; i < EV; ++i) dxs_acc[v][i] += stored<T>(dz[i]);
        }
      }
    }
  }
  // CTA-level reduction of the per-lane-group column sums (fixed order), then one partial row per CTA
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (v * LPR + l) * EV;
#pragma unroll
    for (int i = 0; i < EV; ++i) {
      red[warp * RPI + sub][0][col + i] = da_acc[v][i];
      red[warp * RPI + sub][1][col + i] = db_acc[v][i];
      red[warp * RPI + sub][2][col + i] = dxs_acc[v][i];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 3 * D; e += kLnWarps * 32) {
    const int which = e / D, col = e % D;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kLnWarps * RPI; ++w) s += red[w][which][col];
    dab_ws[((long long)blockIdx.x * 3 + which) * D + col] = s;      // per CTA: [da | db | colsum(dx)]
  }
}

// da[col] / db[col] = sum over the per-CTA partial rows: a CTA owns 32 columns of one of the two vectors, its 8 warps
// take interleaved partial rows (coalesced 128 B reads), fixed-order combine through shared memory (deterministic)
__global__ void __launch_bounds__(256)
ln_dab_finish_kernel(const float* __restrict__ ws, float* __restrict__ da, float* __restrict__ db, int nblk, int D) {
  pdl_wait();
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int cols32 = (D + 31) / 32;
  const int which = blockIdx.x / cols32, col = (blockIdx.x % cols32) * 32 + lane;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  if (col < D) {
    int b = w;
    for (; b + 24 < nblk; b += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) s[u] += ws[((long long)(b + 8 * u) * 3 + which) * D + col];
    }
    for (; b < nblk; b += 8) s[0] += ws[((long long)b * 3 + which) * D + col];
  }
  red[w][lane] = (s[0] + s[1]) + (s[2] + s[3]);
  __syncthreads();
  if (w == 0 && col < D) {
    float t = red[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += red[k][lane];
    (which == 0 ? da : db)[col] = t;               // overwritten: the caller hands in the gradient slot itself
  }
}

// (EV, LPR, VPL) from the row width: the narrowest lane group whose capacity LPR*VPL*EV covers D
template <typename T>
static int fwd_t(const void* x, const void* res, const float* a, const float* b, void* y, float* mean, float* rinv,
                 int rows, int D, float eps, const pka_dropout& dr, cudaStream_t st) {
  dim3 block(kLnWarps * 32);
  const bool drop = dr.p > 0.f;
  // persistent grid: at most 4 resident CTAs per SM (the 8-elements-per-lane instantiations), one iteration per warp for
  // the small in-step tensors
#define LN_FWD(E, LP, V)                                                                                              \
  do {                                                                                                                \
    const long long per_cta = (long long)kLnWarps * (32 / LP);                                                        \
    long long g = (rows + per_cta - 1) / per_cta;                                                                     \
    if (g > (long long)kNumSMs * 4) g = (long long)kNumSMs * 4;                                                       \
    if (drop) launch_k(add_ln_fwd_kernel<T, E, LP, V, true>, dim3((unsigned)g), block, 0, st, (const T*)x, (const T*)res, a, b, (T*)y, mean, rinv, rows, D, eps, dr); \
    else launch_k(add_ln_fwd_kernel<T, E, LP, V, false>, dim3((unsigned)g), block, 0, st, (const T*)x, (const T*)res, a, b, (T*)y, mean, rinv, rows, D, eps, dr); \
  } while (0)
  if (sizeof(T) == 2 && D % 8 == 0) {              // bf16: 8 elements (16 bytes) per vector
    constexpr int E = sizeof(T) == 2 ? 8 : 4;
    const int nv = D / 8;
    if (nv <= 8) LN_FWD(E, 8, 1); else if (nv <= 16) LN_FWD(E, 16, 1); else if (nv <= 32) LN_FWD(E, 32, 1);
    else if (nv <= 64) LN_FWD(E, 32, 2); else LN_FWD(E, 32, 4);
  } else {                                         // fp32 (16 bytes) / bf16 rows that are only a multiple of 4 (8 bytes)
    const int nv = D / 4;
    if (nv <= 8) LN_FWD(4, 8, 1); else if (nv <= 16) LN_FWD(4, 16, 1); else if (nv <= 32) LN_FWD(4, 32, 1);
    else if (nv <= 64) LN_FWD(4, 32, 2); else if (nv <= 128) LN_FWD(4, 32, 4); else LN_FWD(4, 32, 8);
  }
#undef LN_FWD
  return check_launch("add_layernorm_fwd");
}

template <typename T>
static int bwd_t(const void* dy, const void* x, const void* res, const float* a, const float* mean, const float* rinv,
                 void* dx, void* dres, float* da, float* db, float* ws, int rows, int D, float eps,
                 const pka_dropout& dr, cudaStream_t st) {
  const int nblk = pka_ln_bwd_blocks(rows);
  dim3 grid(nblk), block(kLnWarps * 32);
  const bool drop = dr.p > 0.f;
#define LN_BWD(E, LP, V)                                                                                              \
  do {                                                                                                                \
    if (drop) launch_k(add_ln_bwd_kernel<T, E, LP, V, true>, grid, block, 0, st, (const T*)dy, (const T*)x, (const T*)res, a, mean, rinv, (T*)dx, (T*)dres, ws, rows, D, eps, dr); \
    else launch_k(add_ln_bwd_kernel<T, E, LP, V, false>, grid, block, 0, st, (const T*)dy, (const T*)x, (const T*)res, a, mean, rinv, (T*)dx, (T*)dres, ws, rows, D, eps, dr); \
  } while (0)
  if (sizeof(T) == 2 && D % 8 == 0) {
    constexpr int E = sizeof(T) == 2 ? 8 : 4;
    const int nv = D / 8;
    if (nv <= 8) LN_BWD(E, 8, 1); else if (nv <= 16) LN_BWD(E, 16, 1); else if (nv <= 32) LN_BWD(E, 32, 1); else LN_BWD(E, 32, 2);
  } else {
    const int nv = D / 4;
    if (nv <= 8) LN_BWD(4, 8, 1); else if (nv <= 16) LN_BWD(4, 16, 1); else if (nv <= 32) LN_BWD(4, 32, 1);
    else if (nv <= 64) LN_BWD(4, 32, 2); else LN_BWD(4, 32, 4);
  }
#undef LN_BWD
  int rc = check_launch("add_layernorm_bwd");
  if (rc || (!da && !db)) return rc;               // no da/db: the caller finishes the partial rows (pka_reduce_jobs)
  launch_k(ln_dab_finish_kernel, 2 * ((D + 31) / 32), 256, 0, st, ws, da, db, nblk, D);
  return check_launch("ln_dab_finish");
}

}  // namespace pka

extern "C" int pka_ln_bwd_blocks(int rows) {
  // Partial rows (= CTAs) of the backward kernel.  Small decoder tensors: one row per warp (a warp that walks several
  // rows pays one dependent round trip to L2 per row), at most 256 partial rows for the batched reduction launch.
  // Large tensors: a persistent grid that is a whole number of waves for 3, 2 and 1 resident CTAs per SM (the register
  // budgets of the 8- and 16-elements-per-lane instantiations); mid-sized ones (a few rows per warp) get one CTA pair
  // per SM so that every warp still walks several rows behind its prefetch.
  int need = (rows + pka::kLnWarps - 1) / pka::kLnWarps;
  if (need < 1) need = 1;
  if (rows <= 8192) return need < 256 ? need : 256;
  if (rows < pka::kNumSMs * 6 * pka::kLnWarps * 4) return pka::kNumSMs * 2;
  return pka::kNumSMs * 6;
}

extern "C" int pka_add_layernorm_fwd(const void* x, const void* residual, const float* a, const float* b, void* y,
                                     float* mean, float* rinv, int dtype, int rows, int D, float eps,
                                     const pka_dropout* drop, void* stream) {
  using namespace pka;
  PKA_REQUIRE(x && a && b && y && mean && rinv, PKA_EINVAL, "add_layernorm_fwd: null pointer");
  PKA_REQUIRE(rows > 0 && D >= 4 && D % 4 == 0 && D <= 1024, PKA_EUNSUPPORTED, "add_layernorm_fwd: rows=%d D=%d (need D%%4==0, 4<=D<=1024)", rows, D);
  PKA_REQUIRE(aligned16(a) && aligned16(b) && aligned16(x) && aligned16(y) && (!residual || aligned16(residual)), PKA_EALIGN, "add_layernorm_fwd: pointers must be 16-byte aligned");
  pka_dropout dr = drop ? *drop : no_dropout();
  if (dtype == PKA_F32) return fwd_t<float>(x, residual, a, b, y, mean, rinv, rows, D, eps, dr, as_stream(stream));
  if (dtype == PKA_BF16) return fwd_t<__nv_bfloat16>(x, residual, a, b, y, mean, rinv, rows, D, eps, dr, as_stream(stream));
  PKA_REQUIRE(false, PKA_EUNSUPPORTED, "add_layernorm_fwd: dtype %d", dtype);
}

extern "C" int pka_add_layernorm_bwd(const void* dy, const void* x, const void* residual, const float* a,
                                     const float* mean, const float* rinv, void* dx, void* dres, float* da, float* db,
                                     float* dab_ws, int dtype, int rows, int D, float eps, const pka_dropout* drop,
                                     void* stream) {
  using namespace pka;
  PKA_REQUIRE(dy && x && a && mean && rinv && ((da && db) || (!da && !db)) && dab_ws && (dx || dres), PKA_EINVAL, "add_layernorm_bwd: null pointer");
  PKA_REQUIRE(rows > 0 && D >= 4 && D % 4 == 0 && D <= 512, PKA_EUNSUPPORTED, "add_layernorm_bwd: rows=%d D=%d (need D%%4==0, 4<=D<=512)", rows, D);
  PKA_REQUIRE(aligned16(a) && aligned16(x) && aligned16(dy) && (!dx || aligned16(dx)) && (!dres || aligned16(dres)) && (!residual || aligned16(residual)), PKA_EALIGN, "add_layernorm_bwd: pointers must be 16-byte aligned");
  pka_dropout dr = drop ? *drop : no_dropout();
  if (dtype == PKA_F32) return bwd_t<float>(dy, x, residual, a, mean, rinv, dx, dres, da, db, dab_ws, rows, D, eps, dr, as_stream(stream));
  if (dtype == PKA_BF16) return bwd_t<__nv_bfloat16>(dy, x, residual, a, mean, rinv, dx, dres, da, db, dab_ws, rows, D, eps, dr, as_stream(stream));
  PKA_REQUIRE(false, PKA_EUNSUPPORTED, "add_layernorm_bwd: dtype %d", dtype);
}
