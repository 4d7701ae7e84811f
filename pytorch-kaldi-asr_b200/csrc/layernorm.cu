// Fused [dropout] + residual-add + LayerNormalization (forward / backward), one warp per row, warp-shuffle reductions.
// The normalisation is the reference's own (T/Modules.py:42-51), NOT F.layer_norm:
//      y = (z - mean(z)) / (std_unbiased(z) + eps) * a + b,   z = dropout(x) + residual,  eps = 1e-3 added to sigma.
// HBM-bound: forward moves 3*rows*D elements (x, residual in; y out), backward 5 (dy, x, residual in; dres[, dx] out).
#include "common.cuh"

namespace pka {

constexpr int kLnWarps = 8;

// VPL = float4 vectors per lane; a row of D = up to 128*VPL elements lives in registers.
template <typename T, int VPL>
__global__ void __launch_bounds__(kLnWarps * 32)
add_ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ a,
                  const float* __restrict__ bta, T* __restrict__ y, float* __restrict__ mean_o,
                  float* __restrict__ rinv_o, int rows, int D, float eps, const pka_dropout drop) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + warp;
  if (row >= rows) return;
  DropCtx dc = make_drop(drop);
  const T* xr = x + (long long)row * D;
  const T* rr = res ? res + (long long)row * D : nullptr;
  float4 z[VPL];
  float sum = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c = (v * 32 + lane) * 4;
    z[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < D) {
      float4 xv = ld4(xr + c);
      if (dc.p > 0.f) {
        float4 m = dropout_mul4(dc, ((unsigned long long)row * D + c) >> 2);
        xv.x *= m.x; xv.y *= m.y; xv.z *= m.z; xv.w *= m.w;
      }
      if (rr) { float4 rv = ld4(rr + c); xv.x += rv.x; xv.y += rv.y; xv.z += rv.z; xv.w += rv.w; }
      z[v] = xv;
      sum += (xv.x + xv.y) + (xv.z + xv.w);
    }
  }
  const float mean = warp_sum(sum) / (float)D;
  float sq = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c = (v * 32 + lane) * 4;
    if (c < D) {
      float dx = z[v].x - mean, dy = z[v].y - mean, dz = z[v].z - mean, dw = z[v].w - mean;
      sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  }
  const float sigma = sqrtf(warp_sum(sq) / (float)(D - 1));
  const float rinv = 1.f / (sigma + eps);
  T* yr = y + (long long)row * D;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c = (v * 32 + lane) * 4;
    if (c < D) {
      const float4 av = *reinterpret_cast<const float4*>(a + c);
      const float4 bv = *reinterpret_cast<const float4*>(bta + c);
      float4 o;
      o.x = (z[v].x - mean) * rinv * av.x + bv.x;
      o.y = (z[v].y - mean) * rinv * av.y + bv.y;
      o.z = (z[v].z - mean) * rinv * av.z + bv.z;
      o.w = (z[v].w - mean) * rinv * av.w + bv.w;
      st4(yr + c, o);
    }
  }
  if (lane == 0) { mean_o[row] = mean; rinv_o[row] = rinv; }
}

// backward: dz_i = rinv*(g_i - mean(g)) - c_i * rinv^2 * sum(g*c) / ((D-1)*sigma),  g = dy*a, c = z-mean
//           dres = dz,  dx = keep*dz/(1-p),  da += sum_rows dy*c*rinv,  db += sum_rows dy
template <typename T, int VPL>
__global__ void __launch_bounds__(kLnWarps * 32)
add_ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ res,
                  const float* __restrict__ a, const float* __restrict__ mean_i, const float* __restrict__ rinv_i,
                  T* __restrict__ dx, T* __restrict__ dres, float* __restrict__ dab_ws, int rows, int D, float eps,
                  const pka_dropout drop) {
  __shared__ float red[kLnWarps][2][128 * VPL];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  DropCtx dc = make_drop(drop);
  float4 da_acc[VPL], db_acc[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) { da_acc[v] = make_float4(0, 0, 0, 0); db_acc[v] = make_float4(0, 0, 0, 0); }

  for (int row = blockIdx.x * kLnWarps + warp; row < rows; row += gridDim.x * kLnWarps) {
    const float mean = mean_i[row], rinv = rinv_i[row];
    const float sigma = 1.f / rinv - eps;
    const T* xr = x + (long long)row * D;
    const T* rr = res ? res + (long long)row * D : nullptr;
    const T* gr = dy + (long long)row * D;
    float4 c[VPL], g[VPL], keep[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int col = (v * 32 + lane) * 4;
      c[v] = make_float4(0, 0, 0, 0); g[v] = c[v]; keep[v] = make_float4(1, 1, 1, 1);
      if (col < D) {
        float4 xv = ld4(xr + col);
        if (dc.p > 0.f) {
          keep[v] = dropout_mul4(dc, ((unsigned long long)row * D + col) >> 2);
          xv.x *= keep[v].x; xv.y *= keep[v].y; xv.z *= keep[v].z; xv.w *= keep[v].w;
        }
        if (rr) { float4 rv = ld4(rr + col); xv.x += rv.x; xv.y += rv.y; xv.z += rv.z; xv.w += rv.w; }
        c[v] = make_float4(xv.x - mean, xv.y - mean, xv.z - mean, xv.w - mean);
        const float4 dyv = ld4(gr + col);
        const float4 av = *reinterpret_cast<const float4*>(a + col);
        g[v] = make_float4(dyv.x * av.x, dyv.y * av.y, dyv.z * av.z, dyv.w * av.w);
        s1 += (g[v].x + g[v].y) + (g[v].z + g[v].w);
        s2 += (g[v].x * c[v].x + g[v].y * c[v].y) + (g[v].z * c[v].z + g[v].w * c[v].w);
        da_acc[v].x += dyv.x * c[v].x * rinv; da_acc[v].y += dyv.y * c[v].y * rinv;
        da_acc[v].z += dyv.z * c[v].z * rinv; da_acc[v].w += dyv.w * c[v].w * rinv;
        db_acc[v].x += dyv.x; db_acc[v].y += dyv.y; db_acc[v].z += dyv.z; db_acc[v].w += dyv.w;
      }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    const float gm = s1 / (float)D;
    const float kf = sigma > 0.f ? rinv * rinv * s2 / ((float)(D - 1) * sigma) : 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int col = (v * 32 + lane) * 4;
      if (col < D) {
        float4 dz;
        dz.x = rinv * (g[v].x - gm) - c[v].x * kf;
        dz.y = rinv * (g[v].y - gm) - c[v].y * kf;
        dz.z = rinv * (g[v].z - gm) - c[v].z * kf;
        dz.w = rinv * (g[v].w - gm) - c[v].w * kf;
        if (dres) st4(dres + (long long)row * D + col, dz);
        if (dx) {
          dz.x *= keep[v].x; dz.y *= keep[v].y; dz.z *= keep[v].z; dz.w *= keep[v].w;
          st4(dx + (long long)row * D + col, dz);
        }
      }
    }
  }
  // CTA-level reduction of the per-warp column sums, then one partial row per CTA (summed by the finish kernel)
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (v * 32 + lane) * 4;
    *reinterpret_cast<float4*>(&red[warp][0][col]) = da_acc[v];
    *reinterpret_cast<float4*>(&red[warp][1][col]) = db_acc[v];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * D; e += kLnWarps * 32) {
    const int which = e / D, col = e % D;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) s += red[w][which][col];
    dab_ws[((long long)blockIdx.x * 2 + which) * D + col] = s;
  }
}

__global__ void ln_dab_finish_kernel(const float* __restrict__ ws, float* __restrict__ da, float* __restrict__ db,
                                     int nblk, int D) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 2 * D) return;
  const int which = e / D, col = e % D;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  int b = 0;
  for (; b + 4 <= nblk; b += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) s[u] += ws[((long long)(b + u) * 2 + which) * D + col];
  }
  for (; b < nblk; ++b) s[0] += ws[((long long)b * 2 + which) * D + col];
  float* dst = which == 0 ? da : db;
  dst[col] = (s[0] + s[1]) + (s[2] + s[3]);      // overwritten: the caller hands in the gradient slot itself
}

template <typename T>
static int fwd_t(const void* x, const void* res, const float* a, const float* b, void* y, float* mean, float* rinv,
                 int rows, int D, float eps, const pka_dropout& dr, cudaStream_t st) {
  dim3 grid((rows + kLnWarps - 1) / kLnWarps), block(kLnWarps * 32);
  const int vpl = (D + 127) / 128;
#define LN_FWD(V) add_ln_fwd_kernel<T, V><<<grid, block, 0, st>>>((const T*)x, (const T*)res, a, b, (T*)y, mean, rinv, rows, D, eps, dr)
  if (vpl <= 1) LN_FWD(1); else if (vpl <= 2) LN_FWD(2); else if (vpl <= 4) LN_FWD(4); else LN_FWD(8);
#undef LN_FWD
  return check_launch("add_layernorm_fwd");
}

template <typename T>
static int bwd_t(const void* dy, const void* x, const void* res, const float* a, const float* mean, const float* rinv,
                 void* dx, void* dres, float* da, float* db, float* ws, int rows, int D, float eps,
                 const pka_dropout& dr, cudaStream_t st) {
  const int nblk = pka_ln_bwd_blocks(rows);
  dim3 grid(nblk), block(kLnWarps * 32);
  const int vpl = (D + 127) / 128;
#define LN_BWD(V) add_ln_bwd_kernel<T, V><<<grid, block, 0, st>>>((const T*)dy, (const T*)x, (const T*)res, a, mean, rinv, (T*)dx, (T*)dres, ws, rows, D, eps, dr)
  if (vpl <= 1) LN_BWD(1); else if (vpl <= 2) LN_BWD(2); else LN_BWD(4);
#undef LN_BWD
  int rc = check_launch("add_layernorm_bwd");
  if (rc) return rc;
  ln_dab_finish_kernel<<<(2 * D + 255) / 256, 256, 0, st>>>(ws, da, db, nblk, D);
  return check_launch("ln_dab_finish");
}

}  // namespace pka

extern "C" int pka_ln_bwd_blocks(int rows) {
  int need = (rows + pka::kLnWarps - 1) / pka::kLnWarps;
  // few partial rows keep the fixed-order finish short (small decoder tensors); one CTA per SM for large ones
  int cap = rows <= 8192 ? 64 : pka::kNumSMs;
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
  PKA_REQUIRE(dy && x && a && mean && rinv && da && db && dab_ws && (dx || dres), PKA_EINVAL, "add_layernorm_bwd: null pointer");
  PKA_REQUIRE(rows > 0 && D >= 4 && D % 4 == 0 && D <= 512, PKA_EUNSUPPORTED, "add_layernorm_bwd: rows=%d D=%d (need D%%4==0, 4<=D<=512)", rows, D);
  PKA_REQUIRE(aligned16(a) && aligned16(x) && aligned16(dy) && (!dx || aligned16(dx)) && (!dres || aligned16(dres)) && (!residual || aligned16(residual)), PKA_EALIGN, "add_layernorm_bwd: pointers must be 16-byte aligned");
  pka_dropout dr = drop ? *drop : no_dropout();
  if (dtype == PKA_F32) return bwd_t<float>(dy, x, residual, a, mean, rinv, dx, dres, da, db, dab_ws, rows, D, eps, dr, as_stream(stream));
  if (dtype == PKA_BF16) return bwd_t<__nv_bfloat16>(dy, x, residual, a, mean, rinv, dx, dres, da, db, dab_ws, rows, D, eps, dr, as_stream(stream));
  PKA_REQUIRE(false, PKA_EUNSUPPORTED, "add_layernorm_bwd: dtype %d", dtype);
}
