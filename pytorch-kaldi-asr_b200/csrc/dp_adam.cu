// Data-parallel optimiser step as ONE kernel over NVLink peer memory: gradient reduce-scatter + Adam on the rank's
// shard + parameter all-gather, replacing {NCCL all-reduce of the gradient arena, adam_kernel} of the training step.
//
//   every rank holds the whole parameter arena and its own gradient arena in symmetric memory (the same allocation
//   mapped on all ranks: peer pointers, and one NVSwitch multicast address when NVLS is available);
//   rank r owns elements [r*n/W, (r+1)*n/W):
//     A  per-CTA handshake with the same CTA of every peer  -> every rank's backward has finished (its kernel started)
//     1  g = sum over ranks of grad[e]      (multimem.ld_reduce through the switch, or W peer loads in rank order)
//     2  Adam on (param[e], m[e], v[e]) with torch.optim.Adam's arithmetic (the same as adam_kernel)
//     3  param[e] -> every rank             (multimem.st, or W peer stores)
//     B  handshake again (release / acquire at system scope) -> all parameter writes have landed everywhere and nobody
//        still reads this rank's gradients, so the next forward / backward may start
//   Adam moments are only ever touched for the owned shard (ZeRO-1 style: 1/W of the optimiser traffic per rank).
//
// The handshake is the CAS flag protocol (put: 0 -> 1 on the peer's flag, wait: 1 -> 0 on the own flag) on a small
// symmetric flag array laid out [phase][cta][source rank]; flags return to 0, so the kernel is re-launchable and CUDA
// graph capturable.  The grid must be co-resident (<= 148 CTAs) and identical on all ranks.
#include "common.cuh"

namespace pka {

constexpr int kDpMaxWorld = 16;

__device__ __forceinline__ uint32_t cas_relaxed_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.relaxed.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.release.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.acquire.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}

// Thread t < world talks to peer t: raise my flag in the peer's array, then consume the peer's flag in mine.
template <bool ORDERED>
__device__ __forceinline__ void peer_handshake(uint32_t* const* flag_ptrs, int phase, int rank, int world) {
  if ((int)threadIdx.x < world) {
    const int peer = threadIdx.x;
    const size_t slot = ((size_t)phase * gridDim.x + blockIdx.x) * world;
    uint32_t* theirs = flag_ptrs[peer] + slot + rank;
    uint32_t* mine = flag_ptrs[rank] + slot + peer;
    if (ORDERED) {
      while (cas_release_sys(theirs, 0u, 1u) != 0u) {}
      while (cas_acquire_sys(mine, 1u, 0u) != 1u) {}
    } else {
      while (cas_relaxed_sys(theirs, 0u, 1u) != 0u) {}
      while (cas_relaxed_sys(mine, 1u, 0u) != 1u) {}
    }
  }
}

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc_addr) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc_addr) : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st(float* mc_addr, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct DpPeers {
  float* param[kDpMaxWorld];          // parameter arena of every rank (index = rank)
  const float* grad[kDpMaxWorld];     // gradient arena of every rank
};

template <bool MULTICAST>
__global__ void __launch_bounds__(256)
dp_adam_kernel(DpPeers peers, float* param_mc, const float* grad_mc, uint32_t* const* __restrict__ flag_ptrs, int rank,
               int world, float* __restrict__ m, float* __restrict__ v, long long n, const float* __restrict__ lr_dev,
               float lr_host, const long long* __restrict__ state, float b1, float b2, float eps) {
  pdl_wait();
  __shared__ float s_step, s_isb2;
  if (threadIdx.x == 0) {
    const long long t = state[0] + 1;
    const float lr = lr_dev ? lr_dev[0] : lr_host;
    const double bc1 = 1.0 - pow((double)b1, (double)t);
    const double bc2 = 1.0 - pow((double)b2, (double)t);
    s_step = (float)((double)lr / bc1);
    s_isb2 = (float)(1.0 / sqrt(bc2));
  }
  // A: gradients written by kernels that completed before this one started are visible to peers; a relaxed
  // handshake only has to establish that the peer's kernel has started
  peer_handshake<false>(flag_ptrs, 0, rank, world);
  __syncthreads();
  const float step_size = s_step, inv_sqrt_bc2 = s_isb2;
  auto upd = [&](float gv, float& mv, float& vv, float& pv) {
    mv = b1 * mv + (1.f - b1) * gv;
    vv = b2 * vv + (1.f - b2) * gv * gv;
    const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
    pv = pv - step_size * (mv / denom);
  };
  const long long per = (n >> 2) / world;                    // float4 groups per rank (n % (4*world) == 0)
  const long long lo = per * rank;
  float* p_own = peers.param[rank];
  // A peer load is a ~2-3 us round trip over NVLink: every thread keeps kU float4 groups x W ranks of them in flight
  // (round 1 walked its groups one at a time on 32 CTAs: 27 dependent round trips, 90 us for 7 MB on two GPUs).
  constexpr int kU = 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < per; i0 += kU * stride) {
    float4 gv[kU], mv[kU], vv[kU], pv[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long i = i0 + u * stride;
      gv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < per) {
        if (MULTICAST) gv[u] = multimem_ld_reduce_add(grad_mc + 4 * (lo + i));
      }
    }
    if (!MULTICAST) {
      for (int r = 0; r < world; ++r) {                       // fixed rank order: bit-reproducible
        float4 x[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const long long i = i0 + u * stride;
          x[u] = i < per ? __ldcg(reinterpret_cast<const float4*>(peers.grad[r]) + lo + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) { gv[u].x += x[u].x; gv[u].y += x[u].y; gv[u].z += x[u].z; gv[u].w += x[u].w; }
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long i = i0 + u * stride;
      if (i < per) {
        const long long e = lo + i;
        mv[u] = reinterpret_cast<float4*>(m)[e]; vv[u] = reinterpret_cast<float4*>(v)[e];
        pv[u] = reinterpret_cast<const float4*>(p_own)[e];
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long i = i0 + u * stride;
      if (i < per) {
        const long long e = lo + i;
        upd(gv[u].x, mv[u].x, vv[u].x, pv[u].x); upd(gv[u].y, mv[u].y, vv[u].y, pv[u].y);
        upd(gv[u].z, mv[u].z, vv[u].z, pv[u].z); upd(gv[u].w, mv[u].w, vv[u].w, pv[u].w);
        reinterpret_cast<float4*>(m)[e] = mv[u]; reinterpret_cast<float4*>(v)[e] = vv[u];
        if (MULTICAST) {
          multimem_st(param_mc + 4 * e, pv[u]);
        } else {
          for (int r = 0; r < world; ++r) reinterpret_cast<float4*>(peers.param[r])[e] = pv[u];
        }
      }
    }
  }
  // B: every thread's parameter stores precede the release of the flags
  __threadfence_system();
  __syncthreads();
  peer_handshake<true>(flag_ptrs, 1, rank, world);
  __syncthreads();
}

__global__ void dp_adam_t_inc_kernel(long long* state) {
  pdl_wait();
  state[0] += 1;
}

}  // namespace pka

using namespace pka;

extern "C" int pka_dp_adam_grid(int64_t n, int world, int max_ctas) {
  if (n <= 0 || world <= 0 || max_ctas <= 0) return 0;
  const long long per = (n / 4) / world;
  long long ctas = (per + 256 * 2 - 1) / (256 * 2);           // ~2 float4 groups per thread: latency, not bandwidth, bounds it
  if (ctas < 1) ctas = 1;
  if (ctas > max_ctas) ctas = max_ctas;
  if (ctas > kNumSMs) ctas = kNumSMs;
  return (int)ctas;
}

extern "C" int pka_dp_adam_step(const uint64_t* param_ptrs, const uint64_t* grad_ptrs, uint64_t param_mc, uint64_t grad_mc,
                                const uint64_t* flag_ptrs_dev, int rank, int world, int max_ctas, float* exp_avg,
                                float* exp_avg_sq, int64_t n, const float* lr_dev, float lr_host, int64_t* state,
                                float beta1, float beta2, float eps, void* stream) {
  PKA_REQUIRE(param_ptrs && grad_ptrs && flag_ptrs_dev && exp_avg && exp_avg_sq && state && n > 0, PKA_EINVAL,
              "dp_adam_step: bad arguments");
  PKA_REQUIRE(world >= 1 && world <= kDpMaxWorld && rank >= 0 && rank < world, PKA_EINVAL,
              "dp_adam_step: rank %d / world %d out of range (max %d ranks)", rank, world, kDpMaxWorld);
  PKA_REQUIRE(n % (4LL * world) == 0, PKA_EINVAL, "dp_adam_step: arena length %lld must be a multiple of 4 x world", (long long)n);
  PKA_REQUIRE((param_mc == 0) == (grad_mc == 0), PKA_EINVAL, "dp_adam_step: give both multicast addresses or neither");
  PKA_REQUIRE(aligned16(exp_avg) && aligned16(exp_avg_sq), PKA_EALIGN, "dp_adam_step: the moment arenas must be 16-byte aligned");
  DpPeers peers = {};
  for (int r = 0; r < world; ++r) {
    peers.param[r] = reinterpret_cast<float*>(param_ptrs[r]);
    peers.grad[r] = reinterpret_cast<const float*>(grad_ptrs[r]);
    PKA_REQUIRE(peers.param[r] && peers.grad[r] && aligned16(peers.param[r]) && aligned16(peers.grad[r]), PKA_EALIGN,
                "dp_adam_step: peer arena %d missing or not 16-byte aligned", r);
  }
  const int ctas = pka_dp_adam_grid(n, world, max_ctas);
  PKA_REQUIRE(ctas >= 1, PKA_EINVAL, "dp_adam_step: empty grid");
  cudaStream_t st = as_stream(stream);
  uint32_t* const* flags = reinterpret_cast<uint32_t* const*>(flag_ptrs_dev);
  if (param_mc) {
    launch_k(dp_adam_kernel<true>, ctas, 256, 0, st, peers, reinterpret_cast<float*>(param_mc),
             reinterpret_cast<const float*>(grad_mc), flags, rank, world, exp_avg, exp_avg_sq, (long long)n, lr_dev, lr_host,
             (const long long*)state, beta1, beta2, eps);
  } else {
    launch_k(dp_adam_kernel<false>, ctas, 256, 0, st, peers, (float*)nullptr, (const float*)nullptr, flags, rank, world,
             exp_avg, exp_avg_sq, (long long)n, lr_dev, lr_host, (const long long*)state, beta1, beta2, eps);
  }
  int rc = check_launch("dp_adam");
  if (rc) return rc;
  launch_k(dp_adam_t_inc_kernel, 1, 1, 0, st, (long long*)state);
  return check_launch("dp_adam_t_inc");
}
