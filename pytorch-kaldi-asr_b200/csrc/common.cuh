// Shared device helpers for libpka_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/pka_b200.h"

namespace pka {

// ---------------------------------------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define PKA_REQUIRE(cond, code, ...)            \
  do {                                          \
    if (!(cond)) {                              \
      ::pka::set_error(__VA_ARGS__);            \
      return (code);                            \
    }                                           \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------------------------- launches
// Every kernel of the library is launched with programmatic stream serialisation (PDL): the grid may be scheduled while
// the previous kernel of the stream drains, and blocks in pdl_wait() -- the first statement of every kernel, before any
// global-memory access -- until that kernel has completed and flushed.  Inside the captured CUDA graph of the training
// step this removes most of the ~1-2 us launch gap between the ~290 dependent kernels.  PKA_PDL=0 switches it off.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // a failure is picked up by check_launch()
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef PKA_PDL_EARLY
  // Measured: triggering the dependent grid this early gains ~1 % on the training-step graph but costs ~15 % on the
  // beam-search step graph (hundreds of 2-10 us kernels); without it the dependent is released when this grid exits.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
#endif
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kNumSMs = 148;   // B200

// ---------------------------------------------------------------------------------------------- typed load/store
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements <-> float4 (16 B for fp32, 8 B for bf16); pointers must be aligned accordingly
template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> __device__ __forceinline__ void st4(T* p, float4 v);
template <> __device__ __forceinline__ void st4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

// ---------------------------------------------------------------------------------------------- warp / block reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------- Philox4x32-10 dropout
// One call yields 128 bits for counter (idx8, step, site) under key seed, used as EIGHT 16-bit lanes: element i takes
// lane (i & 7) of the call with idx8 = i >> 3 and is kept iff lane >= floor(p * 2^16)  (p is thereby quantised to
// 2^-16: 0.35 -> 0.349991, a 1.4e-5 relative deviation of the expected value, far below bf16/fp32 training noise).
// Every kernel (forward, backward, pka_dropout_mask) derives keep bits only through the helpers below, so masks agree
// by construction.
struct Philox {
  static __device__ __forceinline__ uint4 run(uint64_t seed, uint32_t site, uint64_t step, uint64_t idx) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32) ^ (uint32_t)(step >> 32), c2 = site, c3 = (uint32_t)step;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
      uint32_t h0 = __umulhi(M0, c0), l0 = M0 * c0;
      uint32_t h1 = __umulhi(M1, c2), l1 = M1 * c2;
      uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

struct DropCtx {
  float p;           // 0 = off
  float scale;       // 1/(1-p)
  uint32_t thresh;   // keep iff 16-bit lane >= thresh  (thresh = floor(p * 2^16))
  uint32_t site;
  uint64_t seed;
  uint64_t step;
};

__device__ __forceinline__ DropCtx make_drop(const pka_dropout& d) {
  DropCtx c;
  c.p = d.p;
  c.scale = d.p > 0.f ? 1.f / (1.f - d.p) : 1.f;
  double t = (double)d.p * 65536.0;
  c.thresh = t >= 65535.0 ? 0xffffu : (uint32_t)t;
  c.site = d.site;
  c.seed = d.seed;
  c.step = d.step_ptr ? *d.step_ptr : 0ull;
  return c;
}

// keep bits of elements 8*idx8 .. 8*idx8+7 (bit k = element 8*idx8 + k)
__device__ __forceinline__ uint32_t dropout_bits8(const DropCtx& c, uint64_t idx8) {
  const uint4 r = Philox::run(c.seed, c.site, c.step, idx8);
  uint32_t b = 0u;
  b |= ((r.x & 0xffffu) >= c.thresh ? 1u : 0u) | ((r.x >> 16) >= c.thresh ? 2u : 0u);
  b |= ((r.y & 0xffffu) >= c.thresh ? 4u : 0u) | ((r.y >> 16) >= c.thresh ? 8u : 0u);
  b |= ((r.z & 0xffffu) >= c.thresh ? 16u : 0u) | ((r.z >> 16) >= c.thresh ? 32u : 0u);
  b |= ((r.w & 0xffffu) >= c.thresh ? 64u : 0u) | ((r.w >> 16) >= c.thresh ? 128u : 0u);
  return b;
}
// keep bit of element i
__device__ __forceinline__ bool dropout_keep(const DropCtx& c, uint64_t i) {
  return (dropout_bits8(c, i >> 3) >> (uint32_t)(i & 7)) & 1u;
}
// keep bits of elements 4*idx4 .. 4*idx4+3 as multipliers (0 or scale)
__device__ __forceinline__ float4 dropout_mul4(const DropCtx& c, uint64_t idx4) {
  const uint32_t b = dropout_bits8(c, idx4 >> 1) >> (uint32_t)((idx4 & 1) * 4);
  return make_float4((b & 1u) ? c.scale : 0.f, (b & 2u) ? c.scale : 0.f, (b & 4u) ? c.scale : 0.f, (b & 8u) ? c.scale : 0.f);
}

static inline pka_dropout no_dropout() {
  pka_dropout d;
  d.p = 0.f; d.site = 0; d.seed = 0; d.step_ptr = nullptr;
  return d;
}

}  // namespace pka
