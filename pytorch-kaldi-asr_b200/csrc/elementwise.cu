// HBM-bound element-wise / gather / reduction pieces of the path, the fused Adam step, and library plumbing.
// All kernels use 128-bit accesses where the shape allows, grid-stride loops sized in multiples of the SM count.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

namespace pka {

// ---------------------------------------------------------------------------------------------- error plumbing
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static unsigned long long g_launches = 0;   // kernels launched by this library (bench.py reports it as gpu_launches)
unsigned long long launch_count() { return g_launches; }
int check_launch(const char* what) {
  __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return PKA_ELAUNCH;
  }
  return PKA_OK;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PKA_PDL"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v != 0;
}

static inline int grid_for(long long work_items, int per_block) {
  long long b = (work_items + per_block - 1) / per_block;
  long long cap = (long long)kNumSMs * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---------------------------------------------------------------------------------------------- embedding + position
template <typename T>
__global__ void embed_pos_fwd_kernel(const long long* __restrict__ tok, const float* __restrict__ emb,
                                     const float* __restrict__ pos, T* __restrict__ out, int B, int L, int D, int V,
                                     const pka_dropout drop) {
  pdl_wait();
  DropCtx dc = make_drop(drop);
  const int d4 = D >> 2;
  const long long total = (long long)B * L * d4;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % d4) * 4;
    const long long r = e / d4;
    const int l = (int)(r % L);
    long long t = tok[r];
    if (t < 0 || t >= V) t = 0;
    const float4 ev = *reinterpret_cast<const float4*>(emb + t * D + c);
    const float4 pv = *reinterpret_cast<const float4*>(pos + (long long)l * D + c);
    float4 o = make_float4(ev.x + pv.x, ev.y + pv.y, ev.z + pv.z, ev.w + pv.w);
    if (dc.p > 0.f) {
      const float4 m = dropout_mul4(dc, (unsigned long long)e);      // element index (r*D + c) >> 2 == e
      o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w;
    }
    st4(out + r * D + c, o);
  }
}

// one CTA per vocabulary row.  The token ids of a chunk are staged in shared memory by all threads (coalesced), warp 0
// compacts the positions holding this token (ballot + popc, ascending order), then every thread owns one column and
// sums the matching rows in that fixed order -> deterministic scatter-add, and the D-wide work is only done for the
// ~n_tok/V matching rows instead of all of them.  (Round 1 let warp 0 read the ids from global memory 32 at a time: 63
// dependent round trips for the 2016 tokens of a TIMIT batch, 21 us.)
constexpr int kEmbChunk = 4096;
template <typename T>
__global__ void embed_bwd_kernel(const long long* __restrict__ tok, const T* __restrict__ dout,
                                 float* __restrict__ demb, long long n_tok, int D, int padding_idx,
                                 const pka_dropout drop) {
  pdl_wait();
  __shared__ int stok[kEmbChunk];
  __shared__ int match[kEmbChunk];
  __shared__ int n_match;
  const int v = blockIdx.x;
  if (v == padding_idx) {                          // nn.Embedding(padding_idx): this row never receives a gradient
    for (int c = threadIdx.x; c < D; c += blockDim.x) demb[(long long)v * D + c] = 0.f;
    return;
  }
  DropCtx dc = make_drop(drop);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long base = 0; base < n_tok; base += kEmbChunk) {
    const int n = (int)((n_tok - base) < kEmbChunk ? (n_tok - base) : kEmbChunk);
    for (int i = threadIdx.x; i < n; i += blockDim.x) stok[i] = (int)tok[base + i];
    __syncthreads();
    if (warp == 0) {
      int cnt = 0;
      for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const bool hit = i < n && stok[i] == v;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) match[cnt + __popc(m & ((1u << lane) - 1u))] = i;
        cnt += __popc(m);
      }
      if (lane == 0) n_match = cnt;
    }
    __syncthreads();
    const int nm = n_match;
    // The column loop is warp-uniform when D is a multiple of 32.  A Philox call covers 8 consecutive elements = the
    // columns of an aligned group of 8 lanes, so with dropout the group evaluates 8 DIFFERENT matching rows at once (lane
    // L the row k0 + (L & 7)) and exchanges the bytes: one call + 8 shuffles per 8 rows instead of 8 calls per lane.
    const bool share = dc.p > 0.f && (D & 31) == 0 && (blockDim.x & 31) == 0;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float acc = 0.f;
      if (share) {
        for (int k0 = 0; k0 < nm; k0 += 8) {
          const int kmine = k0 + (lane & 7);
          uint32_t mine = 0u;
          if (kmine < nm) mine = dropout_bits8(dc, (unsigned long long)((base + match[kmine]) * D + c) >> 3);
          float g[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) g[u] = (k0 + u < nm) ? to_f(dout[(base + match[k0 + u]) * D + c]) : 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint32_t bits = __shfl_sync(0xffffffffu, mine, (lane & ~7) | u);
            acc += ((bits >> (lane & 7)) & 1u) ? g[u] * dc.scale : 0.f;     // rows beyond nm contribute g = 0
          }
        }
      } else {
#pragma unroll 4
        for (int k = 0; k < nm; ++k) {
          const long long r = base + match[k];
          float g = to_f(dout[r * D + c]);
          if (dc.p > 0.f) g = dropout_keep(dc, (unsigned long long)(r * D + c)) ? g * dc.scale : 0.f;
          acc += g;
        }
      }
      if (base == 0) demb[(long long)v * D + c] = acc;          // the first chunk defines the row (no zero fill needed)
      else if (nm) demb[(long long)v * D + c] += acc;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------- + row vector, dropout
template <typename T>
__global__ void add_rowvec_dropout_kernel(const T* __restrict__ x, const float* __restrict__ rowvec,
                                          T* __restrict__ out, long long rows, int D, int period,
                                          const pka_dropout drop) {
  pdl_wait();
  DropCtx dc = make_drop(drop);
  const int d4 = D >> 2;
  const long long total = rows * d4;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % d4) * 4;
    const long long r = e / d4;
    float4 o = ld4(x + r * D + c);
    if (rowvec) {
      const float4 pv = *reinterpret_cast<const float4*>(rowvec + (r % period) * D + c);
      o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w;
    }
    if (dc.p > 0.f) {
      const float4 m = dropout_mul4(dc, (unsigned long long)e);
      o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w;
    }
    st4(out + r * D + c, o);
  }
}

template <typename T>
__global__ void dropout_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, long long n4, const pka_dropout drop) {
  pdl_wait();
  DropCtx dc = make_drop(drop);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
    float4 g = ld4(dy + e * 4);
    if (dc.p > 0.f) {
      const float4 m = dropout_mul4(dc, (unsigned long long)e);
      g.x *= m.x; g.y *= m.y; g.z *= m.z; g.w *= m.w;
    }
    st4(dx + e * 4, g);
  }
}

template <typename T>
__global__ void relu_drop_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dz,
                                     long long n4, float scale) {
  pdl_wait();
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
    const float4 g = ld4(dy + e * 4), yv = ld4(y + e * 4);
    float4 o;
    o.x = yv.x > 0.f ? g.x * scale : 0.f;
    o.y = yv.y > 0.f ? g.y * scale : 0.f;
    o.z = yv.z > 0.f ? g.z * scale : 0.f;
    o.w = yv.w > 0.f ? g.w * scale : 0.f;
    st4(dz + e * 4, o);
  }
}

__global__ void dropout_mask_kernel(uint8_t* __restrict__ keep, long long n, const pka_dropout drop) {
  pdl_wait();
  DropCtx dc = make_drop(drop);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    keep[e] = (dc.p > 0.f) ? (dropout_keep(dc, (unsigned long long)e) ? 1 : 0) : 1;
}

// ---------------------------------------------------------------------------------------------- column sums
constexpr int kColsumRowsPerChunk = 64;
template <typename T>
__global__ void colsum_part_kernel(const T* __restrict__ x, float* __restrict__ part, long long rows, int N, int ld) {
  pdl_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long r0 = (long long)blockIdx.y * kColsumRowsPerChunk;
  const long long r1 = r0 + kColsumRowsPerChunk < rows ? r0 + kColsumRowsPerChunk : rows;
  // 8 independent loads in flight per thread; the partial is still a fixed function of the data (deterministic)
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  long long r = r0;
  for (; r + 8 <= r1; r += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) s[u] += to_f(x[(r + u) * ld + n]);
  }
  for (; r < r1; ++r) s[0] += to_f(x[r * ld + n]);
  part[(long long)blockIdx.y * N + n] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}
// vectorised variant (N % 4 == 0, 16-byte aligned rows): a CTA reduces a [128 rows][128 columns] block; a warp reads one
// row segment (32 lanes x 4 columns = 512 B fp32 / 256 B bf16 per request), 8 warps take interleaved rows with 16
// independent loads in flight each, then the 8 row-lanes are added through shared memory in a fixed order.
constexpr int kColsumVecRows = 128;
template <typename T>
__global__ void __launch_bounds__(256)
colsum_part4_kernel(const T* __restrict__ x, float* __restrict__ part, long long rows, int N, int ld) {
  pdl_wait();
  __shared__ float4 red[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 128 + tx * 4;
  const long long r0 = (long long)blockIdx.y * kColsumVecRows;
  float4 acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < N) {
#pragma unroll
    for (int i = 0; i < kColsumVecRows / 8; ++i) {
      const long long r = r0 + ty + 8 * i;
      if (r < rows) {
        const float4 v = ld4<T>(x + r * ld + n);
        acc[i & 3].x += v.x; acc[i & 3].y += v.y; acc[i & 3].z += v.z; acc[i & 3].w += v.w;
      }
    }
  }
  red[ty][tx] = make_float4((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y),
                            (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z), (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w));
  __syncthreads();
  if (ty == 0 && n < N) {
    float4 t = red[0][tx];
#pragma unroll
    for (int w = 1; w < 8; ++w) { t.x += red[w][tx].x; t.y += red[w][tx].y; t.z += red[w][tx].z; t.w += red[w][tx].w; }
    *reinterpret_cast<float4*>(part + (long long)blockIdx.y * N + n) = t;
  }
}
// Fused backward of [ReLU ->] dropout + bias gradient: dZ = gate ? ((Y > 0) ? dY*scale : 0) : dY (bf16) and the column
// sums of dZ in ONE pass over dY / Y (same [128 rows][128 columns] blocking and fixed-order reduction as above).
template <typename Tin>
__global__ void __launch_bounds__(256)
gate_colsum_kernel(const Tin* __restrict__ dY, const __nv_bfloat16* __restrict__ Y, __nv_bfloat16* __restrict__ dZ,
                   float* __restrict__ part, long long rows, int N, float scale, int gate) {
  pdl_wait();
  __shared__ float4 red[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 128 + tx * 4;
  const long long r0 = (long long)blockIdx.y * kColsumVecRows;
  float4 acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < N) {
#pragma unroll
    for (int i = 0; i < kColsumVecRows / 8; ++i) {
      const long long r = r0 + ty + 8 * i;
      if (r < rows) {
        float4 v = ld4<Tin>(dY + r * N + n);
        if (gate) {
          const float4 y = ld4<__nv_bfloat16>(Y + r * N + n);
          v.x = y.x > 0.f ? v.x * scale : 0.f; v.y = y.y > 0.f ? v.y * scale : 0.f;
          v.z = y.z > 0.f ? v.z * scale : 0.f; v.w = y.w > 0.f ? v.w * scale : 0.f;
        }
        st4<__nv_bfloat16>(dZ + r * N + n, v);
        // sum exactly what the GEMMs will read (the bf16-rounded values)
        acc[i & 3].x += __bfloat162float(__float2bfloat16_rn(v.x)); acc[i & 3].y += __bfloat162float(__float2bfloat16_rn(v.y));
        acc[i & 3].z += __bfloat162float(__float2bfloat16_rn(v.z)); acc[i & 3].w += __bfloat162float(__float2bfloat16_rn(v.w));
      }
    }
  }
  red[ty][tx] = make_float4((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y),
                            (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z), (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w));
  __syncthreads();
  if (ty == 0 && n < N) {
    float4 t = red[0][tx];
#pragma unroll
    for (int w = 1; w < 8; ++w) { t.x += red[w][tx].x; t.y += red[w][tx].y; t.z += red[w][tx].z; t.w += red[w][tx].w; }
    *reinterpret_cast<float4*>(part + (long long)blockIdx.y * N + n) = t;
  }
}
// bf16 in / bf16 out variant with 16-byte accesses: a CTA covers [64 rows][256 columns], lane = 8 columns, 8 warps take
// interleaved rows with all eight 16-byte loads of dY (and of Y) in flight before the first use.
__global__ void __launch_bounds__(256)
gate_colsum8_kernel(const __nv_bfloat16* __restrict__ dY, const __nv_bfloat16* __restrict__ Y, __nv_bfloat16* __restrict__ dZ,
                    float* __restrict__ part, long long rows, int N, float scale, int gate) {
  pdl_wait();
  __shared__ float red[8][256];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 256 + tx * 8;
  const long long r0 = (long long)blockIdx.y * 64;
  float acc[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) acc[u] = 0.f;
  if (n < N) {
    uint4 g[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long r = r0 + ty + 8 * i;
      g[i] = make_uint4(0u, 0u, 0u, 0u); y[i] = g[i];
      if (r < rows) {
        g[i] = *reinterpret_cast<const uint4*>(dY + r * N + n);
        if (gate) y[i] = *reinterpret_cast<const uint4*>(Y + r * N + n);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long r = r0 + ty + 8 * i;
      if (r < rows) {
        const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&g[i]);
        const __nv_bfloat162* yh = reinterpret_cast<const __nv_bfloat162*>(&y[i]);
        uint4 o;
        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 v = __bfloat1622float2(gh[k]);
          if (gate) {
            const float2 yv = __bfloat1622float2(yh[k]);
            v.x = yv.x > 0.f ? v.x * scale : 0.f;
            v.y = yv.y > 0.f ? v.y * scale : 0.f;
          }
          oh[k] = __floats2bfloat162_rn(v.x, v.y);
          const float2 w = __bfloat1622float2(oh[k]);          // sum exactly what the GEMMs will read
          acc[2 * k] += w.x; acc[2 * k + 1] += w.y;
        }
        *reinterpret_cast<uint4*>(dZ + r * N + n) = o;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) red[ty][tx * 8 + u] = acc[u];
  __syncthreads();
  const int c = threadIdx.x;                       // 256 threads = 256 columns
  if (blockIdx.x * 256 + c < N) {
    float t = red[0][c];
#pragma unroll
    for (int w = 1; w < 8; ++w) t += red[w][c];
    part[(long long)blockIdx.y * N + blockIdx.x * 256 + c] = t;
  }
}
// out[n] (+)= sum_c part[c][n]: a CTA owns 32 columns, its 8 warps take interleaved chunks (coalesced 128 B rows of the
// partial matrix), fixed-order combine through shared memory
__global__ void __launch_bounds__(256)
colsum_finish_kernel(const float* __restrict__ part, float* __restrict__ out, int chunks, int N, int accumulate) {
  pdl_wait();
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  if (n < N) {
    int c = w;
    for (; c + 24 < chunks; c += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) s[u] += part[(long long)(c + 8 * u) * N + n];
    }
    for (; c < chunks; c += 8) s[0] += part[(long long)c * N + n];
  }
  red[w][lane] = (s[0] + s[1]) + (s[2] + s[3]);
  __syncthreads();
  if (w == 0 && n < N) {
    float t = red[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += red[k][lane];
    out[n] = accumulate ? out[n] + t : t;
  }
}

// ---------------------------------------------------------------------------------------------- cast / transpose
template <typename S, typename Dt>
__global__ void cast_kernel(const S* __restrict__ s, Dt* __restrict__ d, long long n) {
  pdl_wait();
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    d[e] = from_f<Dt>(to_f(s[e]));
}
template <typename S, typename Dt>
__global__ void transpose_kernel(const S* __restrict__ s, Dt* __restrict__ d, int rows, int cols) {
  pdl_wait();
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? to_f(s[(long long)r * cols + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) d[(long long)c * rows + r] = from_f<Dt>(tile[threadIdx.x][i]);
  }
}

// ---------------------------------------------------------------------------------------------- optimiser
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, long long n, const float* lr_dev, float lr_host,
            const long long* state, float b1, float b2, float eps,
            __nv_bfloat16* __restrict__ shadow, unsigned int* done, long long* state_w,
            float* lr_w, int tick_lr, float tick_start_lr, float tick_c) {
  pdl_wait();
  __shared__ float s_step, s_isb2;
  if (threadIdx.x == 0) {
    // bias corrections in double, once per CTA (torch computes them on the host in double precision)
    const long long t = state[0] + 1;
    const float lr = lr_dev ? lr_dev[0] : lr_host;
    const double bc1 = 1.0 - pow((double)b1, (double)t);
    const double bc2 = 1.0 - pow((double)b2, (double)t);
    s_step = (float)((double)lr / bc1);
    s_isb2 = (float)(1.0 / sqrt(bc2));
  }
  __syncthreads();
  const float step_size = s_step, inv_sqrt_bc2 = s_isb2;
  auto upd = [&](float gv, float& mv, float& vv, float& pv) {
    mv = b1 * mv + (1.f - b1) * gv;
    vv = b2 * vv + (1.f - b2) * gv * gv;
    const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
    pv = pv - step_size * (mv / denom);
  };
  const long long n4 = n >> 2;                     // the arena is 16-byte aligned and padded to a multiple of 4
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
    const float4 gv = reinterpret_cast<const float4*>(g)[e];
    float4 mv = reinterpret_cast<float4*>(m)[e], vv = reinterpret_cast<float4*>(v)[e], pv = reinterpret_cast<float4*>(p)[e];
    upd(gv.x, mv.x, vv.x, pv.x); upd(gv.y, mv.y, vv.y, pv.y); upd(gv.z, mv.z, vv.z, pv.z); upd(gv.w, mv.w, vv.w, pv.w);
    reinterpret_cast<float4*>(m)[e] = mv; reinterpret_cast<float4*>(v)[e] = vv; reinterpret_cast<float4*>(p)[e] = pv;
    if (shadow) st4(shadow + 4 * e, pv);
  }
  for (long long e = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float mv = m[e], vv = v[e], pv = p[e];
    upd(g[e], mv, vv, pv);
    m[e] = mv; v[e] = vv; p[e] = pv;
    if (shadow) shadow[e] = __float2bfloat16_rn(pv);
  }
  // The counters advance inside this launch: every CTA read adam_t / lr at its start, so the LAST CTA to finish may
  // store adam_t + 1 (and, when asked, tick the LR schedule: n += 1; lr = start_lr*c/(n+c), T/Optim.py:21-27) without a
  // race -- two single-thread kernels less per step.  `done` is a zero-initialised counter that resets itself.
  if (done) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int prev = atomicAdd(done, 1u);
      if (prev == gridDim.x - 1) {
        state_w[0] += 1;
        if (tick_lr) {
          state_w[1] += 1;
          lr_w[0] = (tick_start_lr * tick_c) / ((float)state_w[1] + tick_c);
        }
        *done = 0u;
        __threadfence();
      }
    }
  }
}
__global__ void adam_t_inc_kernel(long long* state) {
  pdl_wait(); state[0] += 1; }
__global__ void lr_tick_kernel(float* lr, long long* state, float start_lr, float c) {
  pdl_wait();
  state[1] += 1;
  lr[0] = (start_lr * c) / ((float)state[1] + c);
}
__global__ void counter_inc_kernel(unsigned long long* c) {
  pdl_wait(); c[0] += 1ull; }

}  // namespace pka

using namespace pka;

extern "C" const char* pka_last_error(void) { return pka::g_err; }
extern "C" int pka_version(void) { return 100; }
extern "C" uint64_t pka_launch_count(void) { return pka::launch_count(); }
extern "C" int pka_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  PKA_REQUIRE(e == cudaSuccess, PKA_EDEVICE, "check_device: %s", cudaGetErrorString(e));
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  PKA_REQUIRE(e == cudaSuccess, PKA_EDEVICE, "check_device: %s", cudaGetErrorString(e));
  PKA_REQUIRE(major == 10, PKA_EDEVICE, "check_device: compute capability %d.x, this library is sm_100a only", major);
  return PKA_OK;
}

#define DISPATCH_T(dtype, who, CALL)                                         \
  if ((dtype) == PKA_F32) { using T = float; CALL; }                         \
  else if ((dtype) == PKA_BF16) { using T = __nv_bfloat16; CALL; }           \
  else PKA_REQUIRE(false, PKA_EUNSUPPORTED, who ": dtype %d", (int)(dtype))

extern "C" int pka_embed_pos_fwd(const int64_t* tok, const float* emb, const float* pos, void* out, int dtype, int B,
                                 int L, int D, int V, const pka_dropout* drop, void* stream) {
  PKA_REQUIRE(tok && emb && pos && out, PKA_EINVAL, "embed_pos_fwd: null pointer");
  PKA_REQUIRE(B > 0 && L > 0 && D > 0 && D % 4 == 0, PKA_EUNSUPPORTED, "embed_pos_fwd: B=%d L=%d D=%d (D%%4)", B, L, D);
  pka_dropout dr = drop ? *drop : no_dropout();
  const long long total = (long long)B * L * (D / 4);
  DISPATCH_T(dtype, "embed_pos_fwd", (launch_k(embed_pos_fwd_kernel<T>, grid_for(total, 256), 256, 0, as_stream(stream), (const long long*)tok, emb, pos, (T*)out, B, L, D, V, dr)));
  return check_launch("embed_pos_fwd");
}

extern "C" int pka_embed_bwd(const int64_t* tok, const void* dout, float* demb, int dtype, int B, int L, int D, int V,
                             int padding_idx, const pka_dropout* drop, void* stream) {
  PKA_REQUIRE(tok && dout && demb, PKA_EINVAL, "embed_bwd: null pointer");
  PKA_REQUIRE(B > 0 && L > 0 && D > 0 && V > 0, PKA_EINVAL, "embed_bwd: bad sizes");
  pka_dropout dr = drop ? *drop : no_dropout();
  const int threads = D >= 256 ? 256 : (D >= 128 ? 128 : 64);
  DISPATCH_T(dtype, "embed_bwd", (launch_k(embed_bwd_kernel<T>, V, threads, 0, as_stream(stream), (const long long*)tok, (const T*)dout, demb, (long long)B * L, D, padding_idx, dr)));
  return check_launch("embed_bwd");
}

extern "C" int pka_add_rowvec_dropout_fwd(const void* x, const float* rowvec, void* out, int dtype, int64_t rows, int D,
                                          int period, const pka_dropout* drop, void* stream) {
  PKA_REQUIRE(x && out, PKA_EINVAL, "add_rowvec_dropout_fwd: null pointer");
  PKA_REQUIRE(rows > 0 && D > 0 && D % 4 == 0 && (!rowvec || period > 0), PKA_EUNSUPPORTED, "add_rowvec_dropout_fwd: rows=%lld D=%d", (long long)rows, D);
  pka_dropout dr = drop ? *drop : no_dropout();
  const long long total = rows * (D / 4);
  DISPATCH_T(dtype, "add_rowvec_dropout_fwd", (launch_k(add_rowvec_dropout_kernel<T>, grid_for(total, 256), 256, 0, as_stream(stream), (const T*)x, rowvec, (T*)out, rows, D, period, dr)));
  return check_launch("add_rowvec_dropout_fwd");
}

extern "C" int pka_dropout_bwd(const void* dy, void* dx, int dtype, int64_t n, const pka_dropout* drop, void* stream) {
  PKA_REQUIRE(dy && dx, PKA_EINVAL, "dropout_bwd: null pointer");
  PKA_REQUIRE(n > 0 && n % 4 == 0, PKA_EUNSUPPORTED, "dropout_bwd: n=%lld must be a positive multiple of 4", (long long)n);
  pka_dropout dr = drop ? *drop : no_dropout();
  DISPATCH_T(dtype, "dropout_bwd", (launch_k(dropout_bwd_kernel<T>, grid_for(n / 4, 256), 256, 0, as_stream(stream), (const T*)dy, (T*)dx, n / 4, dr)));
  return check_launch("dropout_bwd");
}

extern "C" int pka_relu_drop_bwd(const void* dy, const void* y, void* dz, int dtype, int64_t n, float scale, void* stream) {
  PKA_REQUIRE(dy && y && dz, PKA_EINVAL, "relu_drop_bwd: null pointer");
  PKA_REQUIRE(n > 0 && n % 4 == 0, PKA_EUNSUPPORTED, "relu_drop_bwd: n=%lld must be a positive multiple of 4", (long long)n);
  DISPATCH_T(dtype, "relu_drop_bwd", (launch_k(relu_drop_bwd_kernel<T>, grid_for(n / 4, 256), 256, 0, as_stream(stream), (const T*)dy, (const T*)y, (T*)dz, n / 4, scale)));
  return check_launch("relu_drop_bwd");
}

extern "C" int pka_colsum_chunks(int64_t rows) { return (int)((rows + kColsumRowsPerChunk - 1) / kColsumRowsPerChunk); }

// number of partial rows pka_colsum / pka_gate_colsum write into part_ws for these arguments (for callers that finish
// the sum themselves)
extern "C" int pka_colsum_parts(const void* x, const float* part_ws, int dtype, int64_t rows, int N, int ld) {
  const bool vec = N % 4 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & (dtype == PKA_F32 ? 15u : 7u)) == 0 &&
                   aligned16(part_ws);
  return (int)(vec ? (rows + kColsumVecRows - 1) / kColsumVecRows : (rows + kColsumRowsPerChunk - 1) / kColsumRowsPerChunk);
}
extern "C" int pka_gate_colsum_parts(int dy_dtype, int64_t rows, int N) {
  return (int)((dy_dtype == PKA_BF16 && N % 8 == 0) ? (rows + 63) / 64 : (rows + kColsumVecRows - 1) / kColsumVecRows);
}

extern "C" int pka_colsum(const void* x, float* out, float* part_ws, int dtype, int64_t rows, int N, int ld,
                          int accumulate, void* stream) {
  PKA_REQUIRE(x && part_ws, PKA_EINVAL, "colsum: null pointer");
  PKA_REQUIRE(rows > 0 && N > 0 && ld >= N, PKA_EINVAL, "colsum: rows=%lld N=%d ld=%d", (long long)rows, N, ld);
  int chunks = pka_colsum_chunks(rows);
  PKA_REQUIRE(chunks <= 65535, PKA_EUNSUPPORTED, "colsum: too many rows");
  const bool vec = N % 4 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & (dtype == PKA_F32 ? 15u : 7u)) == 0 &&
                   aligned16(part_ws);
  if (vec) {
    chunks = (int)((rows + kColsumVecRows - 1) / kColsumVecRows);        // never more than the workspace was sized for
    dim3 grid((N + 127) / 128, chunks);
    DISPATCH_T(dtype, "colsum", (launch_k(colsum_part4_kernel<T>, grid, 256, 0, as_stream(stream), (const T*)x, part_ws, rows, N, ld)));
  } else {
    dim3 grid((N + 127) / 128, chunks);
    DISPATCH_T(dtype, "colsum", (launch_k(colsum_part_kernel<T>, grid, 128, 0, as_stream(stream), (const T*)x, part_ws, rows, N, ld)));
  }
  int rc = check_launch("colsum_part");
  if (rc || !out) return rc;                       // no out: the caller sums the partial rows (pka_reduce_jobs)
  launch_k(colsum_finish_kernel, (N + 31) / 32, 256, 0, as_stream(stream), part_ws, out, chunks, N, accumulate);
  return check_launch("colsum_finish");
}

extern "C" int pka_gate_colsum(const void* dY, int dy_dtype, const void* Y, void* dZ, float* out, float* part_ws, int64_t rows,
                               int N, float scale, int gate, void* stream) {
  PKA_REQUIRE(dY && dZ && part_ws && (!gate || Y), PKA_EINVAL, "gate_colsum: null pointer");
  PKA_REQUIRE(rows > 0 && N > 0 && N % 4 == 0, PKA_EUNSUPPORTED, "gate_colsum: rows=%lld N=%d (need N%%4==0)", (long long)rows, N);
  PKA_REQUIRE(aligned16(dY) && aligned16(dZ) && aligned16(part_ws) && (!Y || aligned16(Y)), PKA_EALIGN, "gate_colsum: pointers must be 16-byte aligned");
  int chunks = (int)((rows + kColsumVecRows - 1) / kColsumVecRows);
  PKA_REQUIRE((rows + 63) / 64 <= 65535, PKA_EUNSUPPORTED, "gate_colsum: too many rows");
  dim3 grid((N + 127) / 128, chunks);
  if (dy_dtype == PKA_BF16 && N % 8 == 0) {
    chunks = (int)((rows + 63) / 64);              // = pka_colsum_chunks(rows): what the workspace is sized for
    launch_k(gate_colsum8_kernel, dim3((N + 255) / 256, chunks), 256, 0, as_stream(stream), (const __nv_bfloat16*)dY,
             (const __nv_bfloat16*)Y, (__nv_bfloat16*)dZ, part_ws, (long long)rows, N, scale, gate);
  } else if (dy_dtype == PKA_BF16)
    launch_k(gate_colsum_kernel<__nv_bfloat16>, grid, 256, 0, as_stream(stream), (const __nv_bfloat16*)dY, (const __nv_bfloat16*)Y, (__nv_bfloat16*)dZ, part_ws, (long long)rows, N, scale, gate);
  else if (dy_dtype == PKA_F32)
    launch_k(gate_colsum_kernel<float>, grid, 256, 0, as_stream(stream), (const float*)dY, (const __nv_bfloat16*)Y, (__nv_bfloat16*)dZ, part_ws, (long long)rows, N, scale, gate);
  else PKA_REQUIRE(false, PKA_EUNSUPPORTED, "gate_colsum: dtype %d", dy_dtype);
  int rc = check_launch("gate_colsum");
  if (rc || !out) return rc;
  launch_k(colsum_finish_kernel, (N + 31) / 32, 256, 0, as_stream(stream), part_ws, out, chunks, N, 0);
  return check_launch("colsum_finish");
}

extern "C" int pka_dropout_mask(uint8_t* keep, int64_t n, const pka_dropout* drop, void* stream) {
  PKA_REQUIRE(keep && drop && n > 0, PKA_EINVAL, "dropout_mask: bad arguments");
  launch_k(dropout_mask_kernel, grid_for(n, 256), 256, 0, as_stream(stream), keep, n, *drop);
  return check_launch("dropout_mask");
}

extern "C" int pka_cast(const void* src, int sd, void* dst, int dd, int64_t n, void* stream) {
  PKA_REQUIRE(src && dst && n > 0, PKA_EINVAL, "cast: bad arguments");
  cudaStream_t st = as_stream(stream);
  const int g = grid_for(n, 256);
  if (sd == PKA_F32 && dd == PKA_BF16) launch_k(cast_kernel<float, __nv_bfloat16>, g, 256, 0, st, (const float*)src, (__nv_bfloat16*)dst, n);
  else if (sd == PKA_BF16 && dd == PKA_F32) launch_k(cast_kernel<__nv_bfloat16, float>, g, 256, 0, st, (const __nv_bfloat16*)src, (float*)dst, n);
  else if (sd == PKA_F32 && dd == PKA_F32) launch_k(cast_kernel<float, float>, g, 256, 0, st, (const float*)src, (float*)dst, n);
  else if (sd == PKA_BF16 && dd == PKA_BF16) launch_k(cast_kernel<__nv_bfloat16, __nv_bfloat16>, g, 256, 0, st, (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n);
  else PKA_REQUIRE(false, PKA_EUNSUPPORTED, "cast: dtypes %d -> %d", sd, dd);
  return check_launch("cast");
}

extern "C" int pka_transpose(const void* src, int sd, void* dst, int dd, int rows, int cols, void* stream) {
  PKA_REQUIRE(src && dst && rows > 0 && cols > 0, PKA_EINVAL, "transpose: bad arguments");
  cudaStream_t st = as_stream(stream);
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  PKA_REQUIRE(grid.y <= 65535, PKA_EUNSUPPORTED, "transpose: too many rows");
  if (sd == PKA_F32 && dd == PKA_BF16) launch_k(transpose_kernel<float, __nv_bfloat16>, grid, block, 0, st, (const float*)src, (__nv_bfloat16*)dst, rows, cols);
  else if (sd == PKA_BF16 && dd == PKA_F32) launch_k(transpose_kernel<__nv_bfloat16, float>, grid, block, 0, st, (const __nv_bfloat16*)src, (float*)dst, rows, cols);
  else if (sd == PKA_F32 && dd == PKA_F32) launch_k(transpose_kernel<float, float>, grid, block, 0, st, (const float*)src, (float*)dst, rows, cols);
  else if (sd == PKA_BF16 && dd == PKA_BF16) launch_k(transpose_kernel<__nv_bfloat16, __nv_bfloat16>, grid, block, 0, st, (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, rows, cols);
  else PKA_REQUIRE(false, PKA_EUNSUPPORTED, "transpose: dtypes %d -> %d", sd, dd);
  return check_launch("transpose");
}

extern "C" int pka_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                             const float* lr_dev, float lr_host, int64_t* state, float beta1, float beta2, float eps,
                             void* bf16_shadow, void* stream) {
  PKA_REQUIRE(param && grad && exp_avg && exp_avg_sq && state && n > 0, PKA_EINVAL, "adam_step: bad arguments");
  cudaStream_t st = as_stream(stream);
  PKA_REQUIRE(aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq), PKA_EALIGN, "adam_step: the arenas must be 16-byte aligned");
  launch_k(adam_kernel, grid_for((n + 3) / 4, 256), 256, 0, st, param, grad, exp_avg, exp_avg_sq, n, lr_dev, lr_host, (const long long*)state, beta1, beta2, eps, (__nv_bfloat16*)bf16_shadow,
           (unsigned int*)nullptr, (long long*)nullptr, (float*)nullptr, 0, 0.f, 0.f);
  int rc = check_launch("adam");
  if (rc) return rc;
  launch_k(adam_t_inc_kernel, 1, 1, 0, st, (long long*)state);
  return check_launch("adam_t_inc");
}

// pka_adam_step with the counter updates folded into the same launch (see adam_kernel): adam_t += 1 always; when
// tick_lr != 0 also the LR schedule tick of pka_lr_tick (lr_dev must then be given).  done_counter: uint32[1], zero
// before the first call, owned by the caller.
extern "C" int pka_adam_step_fused(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                   float* lr_dev, float lr_host, int64_t* state, float beta1, float beta2, float eps,
                                   void* bf16_shadow, uint32_t* done_counter, int tick_lr, float start_lr,
                                   float soft_coefficient, void* stream) {
  PKA_REQUIRE(param && grad && exp_avg && exp_avg_sq && state && done_counter && n > 0 && (!tick_lr || lr_dev), PKA_EINVAL,
              "adam_step_fused: bad arguments");
  PKA_REQUIRE(aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq), PKA_EALIGN, "adam_step_fused: the arenas must be 16-byte aligned");
  launch_k(adam_kernel, grid_for((n + 3) / 4, 256), 256, 0, as_stream(stream), param, grad, exp_avg, exp_avg_sq, n, (const float*)lr_dev, lr_host,
           (const long long*)state, beta1, beta2, eps, (__nv_bfloat16*)bf16_shadow, (unsigned int*)done_counter, (long long*)state, lr_dev,
           tick_lr, start_lr, soft_coefficient);
  return check_launch("adam_fused");
}

extern "C" int pka_lr_tick(float* lr_dev, int64_t* state, float start_lr, float soft_coefficient, void* stream) {
  PKA_REQUIRE(lr_dev && state, PKA_EINVAL, "lr_tick: null pointer");
  launch_k(lr_tick_kernel, 1, 1, 0, as_stream(stream), lr_dev, (long long*)state, start_lr, soft_coefficient);
  return check_launch("lr_tick");
}

extern "C" int pka_counter_inc(uint64_t* counter, void* stream) {
  PKA_REQUIRE(counter, PKA_EINVAL, "counter_inc: null pointer");
  launch_k(counter_inc_kernel, 1, 1, 0, as_stream(stream), (unsigned long long*)counter);
  return check_launch("counter_inc");
}
