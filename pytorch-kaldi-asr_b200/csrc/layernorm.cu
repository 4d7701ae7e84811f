// Fused [dropout] + residual-add + LayerNormalization (forward / backward), one warp per row, warp-shuffle reductions.
// The normalisation is the reference's own (T/Modules.py:42-51), NOT F.layer_norm:
//      y = (z - mean(z)) / (std_unbiased(z) + eps) * a + b,   z = dropout(x) + residual,  eps = 1e-3 added to sigma.
// HBM-bound: forward moves 3*rows*D elements (x, residual in; y out), backward 5 (dy, x, residual in; dres[, dx] out).
#include "common.cuh"

namespace pka {

constexpr int kLnWarps = 8;

// EV consecutive elements (one 16-byte vector: 4 fp32 or 8 bf16) <-> registers
template <typename T, int EV> __device__ __forceinline__ void ldv(const T* p, float (&v)[EV]);
template <> __device__ __forceinline__ void ldv<float, 4>(const float* p, float (&v)[4]) {
  const float4 q = *reinterpret_cast<const float4*>(p);
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
template <> __device__ __forceinline__ void ldv<__nv_bfloat16, 4>(const __nv_bfloat16* p, float (&v)[4]) {
  const float4 q = ld4(p);
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
template <> __device__ __forceinline__ void ldv<__nv_bfloat16, 8>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <typename T, int EV> __device__ __forceinline__ void stv(T* p, const float (&v)[EV]);
template <> __device__ __forceinline__ void stv<float, 4>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void stv<__nv_bfloat16, 4>(__nv_bfloat16* p, const float (&v)[4]) {
  st4(p, make_float4(v[0], v[1], v[2], v[3]));
}
template <> __device__ __forceinline__ void stv<__nv_bfloat16, 8>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}
// dropout multipliers of EV consecutive elements starting at element index e0 (e0 % EV == 0)
template <int EV> __device__ __forceinline__ void drop_mulv(const DropCtx& dc, unsigned long long e0, float (&m)[EV]) {
  if (EV == 8) {
    const uint32_t b = dropout_bits8(dc, e0 >> 3);
#pragma unroll
    for (int i = 0; i < EV; ++i) m[i] = ((b >> i) & 1u) ? dc.scale : 0.f;
  } else {
    const float4 q = dropout_mul4(dc, e0 >> 2);
    m[0] = q.x; m[1] = q.y; m[2] = q.z; m[3] = q.w;
  }
}

// VPL = 16-byte vectors per lane (EV elements each); a row of D <= 32*EV*VPL elements lives in registers.
// RPW = rows per warp: the loads of all RPW rows (x and residual) are issued before the first reduction, so a warp has
// RPW times the bytes in flight (narrow rows -- D = 128 bf16 is 256 B -- cannot cover the HBM latency one row at a time:
// 48 % of the copy bandwidth with RPW = 1 at a bandwidth-sized shape, round 2).
template <typename T, int EV, int VPL, int RPW>
__global__ void __launch_bounds__(kLnWarps * 32)
add_ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ a,
                  const float* __restrict__ bta, T* __restrict__ y, float* __restrict__ mean_o,
                  float* __restrict__ rinv_o, int rows, int D, float eps, const pka_dropout drop) {
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * kLnWarps + warp) * RPW;
  if (row0 >= rows) return;
  DropCtx dc = make_drop(drop);
  float z[RPW][VPL][EV], rv[RPW][VPL][EV];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {                  // every load of the warp's rows first
    const int row = row0 + r;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = (v * 32 + lane) * EV;
#pragma unroll
      for (int i = 0; i < EV; ++i) { z[r][v][i] = 0.f; rv[r][v][i] = 0.f; }
      if (c < D && row < rows) {
        ldv<T, EV>(x + (long long)row * D + c, z[r][v]);
        if (res) ldv<T, EV>(res + (long long)row * D + c, rv[r][v]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int row = row0 + r;
    if (row >= rows) break;                        // warp-uniform
    float sum = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = (v * 32 + lane) * EV;
      if (c < D) {
        if (dc.p > 0.f) {
          float m[EV];
          drop_mulv<EV>(dc, (unsigned long long)row * D + c, m);
#pragma unroll
          for (int i = 0; i < EV; ++i) z[r][v][i] *= m[i];
        }
#pragma unroll
        for (int i = 0; i < EV; ++i) { z[r][v][i] += rv[r][v][i]; sum += z[r][v][i]; }
      }
    }
    const float mean = warp_sum(sum) / (float)D;
    float sq = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = (v * 32 + lane) * EV;
      if (c < D) {
#pragma unroll
        for (int i = 0; i < EV; ++i) { const float d = z[r][v][i] - mean; sq = fmaf(d, d, sq); }
      }
    }
    const float sigma = sqrtf(warp_sum(sq) / (float)(D - 1));
    const float rinv = 1.f / (sigma + eps);
    T* yr = y + (long long)row * D;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = (v * 32 + lane) * EV;
      if (c < D) {
        float o[EV];
#pragma unroll
        for (int i = 0; i < EV; i += 4) {
          const float4 av = *reinterpret_cast<const float4*>(a + c + i);
          const float4 bv = *reinterpret_cast<const float4*>(bta + c + i);
          o[i] = (z[r][v][i] - mean) * rinv * av.x + bv.x;
          o[i + 1] = (z[r][v][i + 1] - mean) * rinv * av.y + bv.y;
          o[i + 2] = (z[r][v][i + 2] - mean) * rinv * av.z + bv.z;
          o[i + 3] = (z[r][v][i + 3] - mean) * rinv * av.w + bv.w;
        }
        stv<T, EV>(yr + c, o);
      }
    }
    if (lane == 0) { mean_o[row] = mean; rinv_o[row] = rinv; }
  }
}

// backward: dz_i = rinv*(g_i - mean(g)) - c_i * rinv^2 * sum(g*c) / ((D-1)*sigma),  g = dy*a, c = z-mean
//           dres = dz,  dx = keep*dz/(1-p),  da = sum_rows dy*c*rinv,  db = sum_rows dy
template <typename T, int EV, int VPL>
__global__ void __launch_bounds__(kLnWarps * 32)
add_ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ res,
                  const float* __restrict__ a, const float* __restrict__ mean_i, const float* __restrict__ rinv_i,
                  T* __restrict__ dx, T* __restrict__ dres, float* __restrict__ dab_ws, int rows, int D, float eps,
                  const pka_dropout drop) {
  pdl_wait();
  __shared__ float red[kLnWarps][3][32 * EV * VPL];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  DropCtx dc = make_drop(drop);
  float da_acc[VPL][EV], db_acc[VPL][EV], dxs_acc[VPL][EV];       // dxs: column sums of the dx this kernel writes
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
#pragma unroll
    for (int i = 0; i < EV; ++i) { da_acc[v][i] = 0.f; db_acc[v][i] = 0.f; dxs_acc[v][i] = 0.f; }
  }

  for (int row = blockIdx.x * kLnWarps + warp; row < rows; row += gridDim.x * kLnWarps) {
    const float mean = mean_i[row], rinv = rinv_i[row];
    const float sigma = 1.f / rinv - eps;
    const T* xr = x + (long long)row * D;
    const T* rr = res ? res + (long long)row * D : nullptr;
    const T* gr = dy + (long long)row * D;
    float c[VPL][EV], g[VPL][EV], keep[VPL][EV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int col = (v * 32 + lane) * EV;
#pragma unroll
      for (int i = 0; i < EV; ++i) { c[v][i] = 0.f; g[v][i] = 0.f; keep[v][i] = 1.f; }
      if (col < D) {
        float xv[EV], dyv[EV];
        ldv<T, EV>(xr + col, xv);
        ldv<T, EV>(gr + col, dyv);
        if (dc.p > 0.f) {
          drop_mulv<EV>(dc, (unsigned long long)row * D + col, keep[v]);
#pragma unroll
          for (int i = 0; i < EV; ++i) xv[i] *= keep[v][i];
        }
        if (rr) {
          float rv[EV];
          ldv<T, EV>(rr + col, rv);
#pragma unroll
          for (int i = 0; i < EV; ++i) xv[i] += rv[i];
        }
        float av[EV];                              // L1-resident gain vector (kept out of the loop-carried registers)
#pragma unroll
        for (int i = 0; i < EV; i += 4) {
          const float4 q = *reinterpret_cast<const float4*>(a + col + i);
          av[i] = q.x; av[i + 1] = q.y; av[i + 2] = q.z; av[i + 3] = q.w;
        }
#pragma unroll
        for (int i = 0; i < EV; ++i) {
          c[v][i] = xv[i] - mean;
          g[v][i] = dyv[i] * av[i];
          s1 += g[v][i];
          s2 = fmaf(g[v][i], c[v][i], s2);
          da_acc[v][i] = fmaf(dyv[i] * c[v][i], rinv, da_acc[v][i]);
          db_acc[v][i] += dyv[i];
        }
      }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    const float gm = s1 / (float)D;
    const float kf = sigma > 0.f ? rinv * rinv * s2 / ((float)(D - 1) * sigma) : 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int col = (v * 32 + lane) * EV;
      if (col < D) {
        float dz[EV];
#pragma unroll
        for (int i = 0; i < EV; ++i) dz[i] = rinv * (g[v][i] - gm) - c[v][i] * kf;
        if (dres) stv<T, EV>(dres + (long long)row * D + col, dz);
        if (dx) {
#pragma unroll
          for (int i = 0; i < EV; ++i) dz[i] *= keep[v][i];
          stv<T, EV>(dx + (long long)row * D + col, dz);
        }
        // bias gradient of the linear layer below = column sums of exactly what it will read (the stored, rounded dx)
#pragma unroll
        for (int i = 0; i < EV; ++i) dxs_acc[v][i] += to_f(from_f<T>(dz[i]));
      }
    }
  }
  // CTA-level reduction of the per-warp column sums, then one partial row per CTA (summed by the finish kernel)
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (v * 32 + lane) * EV;
#pragma unroll
    for (int i = 0; i < EV; ++i) { red[warp][0][col + i] = da_acc[v][i]; red[warp][1][col + i] = db_acc[v][i]; red[warp][2][col + i] = dxs_acc[v][i]; }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 3 * D; e += kLnWarps * 32) {
    const int which = e / D, col = e % D;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) s += red[w][which][col];
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

template <typename T>
static int fwd_t(const void* x, const void* res, const float* a, const float* b, void* y, float* mean, float* rinv,
                 int rows, int D, float eps, const pka_dropout& dr, cudaStream_t st) {
  dim3 block(kLnWarps * 32);
  // rows per warp: 4 for narrow rows of a large tensor (bytes in flight), 1 otherwise (small tensors want every SM busy)
#define LN_FWD(E, V, R) launch_k(add_ln_fwd_kernel<T, E, V, R>, dim3((rows + kLnWarps * R - 1) / (kLnWarps * R)), block, 0, st, (const T*)x, (const T*)res, a, b, (T*)y, mean, rinv, rows, D, eps, dr)
  const bool many = rows >= 16384;
  if (sizeof(T) == 2 && D % 8 == 0 && D > 128) {  // bf16: 8 elements (16 bytes) per lane and vector (D <= 128 would
                                                  // leave half of the warp idle: 4 elements per lane there)
    constexpr int E = sizeof(T) == 2 ? 8 : 4;
    const int vpl = (D + 255) / 256;
    if (vpl <= 1) { if (many) LN_FWD(E, 1, 2); else LN_FWD(E, 1, 1); } else if (vpl <= 2) LN_FWD(E, 2, 1); else LN_FWD(E, 4, 1);
  } else {
    const int vpl = (D + 127) / 128;
    if (vpl <= 1) { if (many) LN_FWD(4, 1, 4); else LN_FWD(4, 1, 1); }
    else if (vpl <= 2) { if (many) LN_FWD(4, 2, 2); else LN_FWD(4, 2, 1); }
    else if (vpl <= 4) LN_FWD(4, 4, 1); else LN_FWD(4, 8, 1);
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
#define LN_BWD(E, V) launch_k(add_ln_bwd_kernel<T, E, V>, grid, block, 0, st, (const T*)dy, (const T*)x, (const T*)res, a, mean, rinv, (T*)dx, (T*)dres, ws, rows, D, eps, dr)
  if (sizeof(T) == 2 && D % 8 == 0 && D > 128) {
    constexpr int E = sizeof(T) == 2 ? 8 : 4;
    const int vpl = (D + 255) / 256;
    if (vpl <= 1) LN_BWD(E, 1); else LN_BWD(E, 2);
  } else {
    const int vpl = (D + 127) / 128;
    if (vpl <= 1) LN_BWD(4, 1); else if (vpl <= 2) LN_BWD(4, 2); else LN_BWD(4, 4);
  }
#undef LN_BWD
  int rc = check_launch("add_layernorm_bwd");
  if (rc || (!da && !db)) return rc;               // no da/db: the caller finishes the partial rows (pka_reduce_jobs)
  launch_k(ln_dab_finish_kernel, 2 * ((D + 31) / 32), 256, 0, st, ws, da, db, nblk, D);
  return check_launch("ln_dab_finish");
}

}  // namespace pka

extern "C" int pka_ln_bwd_blocks(int rows) {
  int need = (rows + pka::kLnWarps - 1) / pka::kLnWarps;
  // few partial rows keep the fixed-order finish short (small decoder tensors); one CTA per SM for large ones
  // one row per warp for the small decoder tensors (a warp walks its rows serially: 4 rows = 4 dependent round trips
  // to L2); the partial rows are summed by the batched reduction launch, so their number no longer costs a kernel
  int cap = rows <= 8192 ? 256 : pka::kNumSMs * 8; // enough resident warps to pull HBM bandwidth on large tensors
  return need < cap ? (need > 0 ? need : 1) : cap;
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
