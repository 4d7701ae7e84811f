// Weight gradient of the bf16 tensor-core path, CTA-pair version (tcgen05.mma cta_group::2):
//
//   dW[o, seg*N + i] = sum_{b,t} dZ[b,t,o] * X[b, t+shift[seg], i]          (pka_gemm_tc mode 2, M % 256 == 0, N % 256 == 0)
//
// Both operands are read straight from the row-major activations as MN-major UMMA tiles (frames are the reduction
// index; the splice shift is a TMA row coordinate, out-of-range frames are zero-filled = ConcatLayer's padding).  A pair
// of CTAs owns a [256 output channels] x [256 input channels] tile of one splice context for one slice of the
// utterances: each CTA stages 128 dZ channels (its half of M) and 128 X channels (its half of N) -- half the
// shared-memory ingress per FLOP of the single-CTA kernel in gemm_tc.cu -- in stages of 128 frames (four 16 KB boxes,
// the largest the 128 B swizzle allows: TMA cost is ~100 ns per box), and the leader issues 256x256x16 MMAs for both SMs.
// fp32 partial tiles go to the split workspace and are summed in fixed order by tc_reduce_kernel (deterministic).
#include "tc_common.cuh"

namespace pka {

constexpr int WG_KF = 128;                                  // frames per stage
constexpr int WG_BOX = WG_KF * 64 * 2;                      // [128 frames][64 channels] bf16 = 16 KB
constexpr int WG_STAGE = 4 * WG_BOX;                        // dZ half (2 boxes) + X half (2 boxes) = 64 KB
constexpr int WG_STAGES = 3;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE + 256;
constexpr int WG_THREADS = 192;

struct WgradParams {
  int Bt, T, M, N, nseg, ldc;
  int shift[PKA_MAX_CTX];
  int units_per_split, n_units, tk_per_utt, n_tiles_n;   // a unit = 128 frames of one utterance
  float* C;
};

__device__ __forceinline__ void wg_tma_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void wg_mma_pair(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void wg_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1)
gemm_tc_wgrad2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                      const __grid_constant__ CUtensorMap mapC, const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = (uint64_t*)(smem + WG_STAGES * WG_STAGE);
  // bars: [0,3) full (leader), [3,6) empty, [6] accumulator ready; then the TMEM slot
  uint32_t* tmem_slot = (uint32_t*)(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool leader = rank == 0;
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  // tile of this pair: (m tile of 256, splice context, n tile of 256); split over utterances along grid.y
  const int tile = blockIdx.x >> 1;
  const int nt = tile % p.n_tiles_n, seg = (tile / p.n_tiles_n) % p.nseg, mt = tile / (p.n_tiles_n * p.nseg);
  const int m0 = mt * 256 + (int)rank * 128;       // this CTA's dZ channels
  const int n0 = nt * 256 + (int)rank * 128;       // this CTA's X channels
  const int u_lo = blockIdx.y * p.units_per_split, u_hi = min(p.n_units, u_lo + p.units_per_split);
  const int n_iters = max(0, u_hi - u_lo);
  const int shift = p.shift[seg];

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[3 + s]), 1); }
    mbar_init(smem_u32(&bars[6]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                      // the prologue above touched only shared / tensor memory

  if (warp == 0) {
    if (lane == 0) {                               // ===== TMA producer (both CTAs; completion lands on the leader's barrier)
      tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapB);
      int s = 0;
      uint32_t ph = 0;
      int b = u_lo / p.tk_per_utt, tk = u_lo - b * p.tk_per_utt;
      for (int u = u_lo; u < u_hi; ++u) {
          mbar_wait(smem_u32(&bars[3 + s]), ph ^ 1u);
          const uint32_t full = smem_u32(&bars[s]);
          if (leader) mbar_expect_tx(full, 2u * WG_STAGE);
          const uint32_t dst = smem_u32(smem + s * WG_STAGE);
          const int t0 = tk * WG_KF;
          wg_tma_pair(dst, &mapA, full, m0, t0, b);
          wg_tma_pair(dst + WG_BOX, &mapA, full, m0 + 64, t0, b);
          wg_tma_pair(dst + 2 * WG_BOX, &mapB, full, n0, t0 + shift, b);
          wg_tma_pair(dst + 3 * WG_BOX, &mapB, full, n0 + 64, t0 + shift, b);
          if (++s == WG_STAGES) { s = 0; ph ^= 1u; }
          if (++tk == p.tk_per_utt) { tk = 0; ++b; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {                     // ===== MMA issuer
      const uint32_t idesc = make_idesc(256, 256, true, true);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < n_iters; ++i) {
        mbar_wait(smem_u32(&bars[s]), ph);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + s * WG_STAGE);
        const uint64_t da = make_sdesc(base, WG_BOX), db = make_sdesc(base + 2 * WG_BOX, WG_BOX);
#pragma unroll
        for (int k = 0; k < WG_KF / 16; ++k)       // 16 frames = two 8-row groups = 2048 B per UMMA_K step
          wg_mma_pair(tmem_base, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc, (i | k) ? 1u : 0u);
        wg_commit_pair(smem_u32(&bars[3 + s]));
        if (++s == WG_STAGES) { s = 0; ph ^= 1u; }
      }
      wg_commit_pair(smem_u32(&bars[6]));
    }
  } else {                                         // ===== epilogue warps 2..5: fp32 partial tile of dW
    // A thread owns one row (output channel); writing its 1 KB straight to global memory would cost 32 scattered
    // 16-byte requests per store instruction.  Instead 32-column slices go through the (now idle) first pipeline stage
    // as 128B-swizzled [128 rows][32 fp32] tiles and leave as bulk tensor stores.
    const int q = warp & 3;
    const int row = q * 32 + lane;
    if (n_iters > 0) {
      mbar_wait(smem_u32(&bars[6]), 0);
      tc_fence_after();
    }
    const int col0 = seg * p.N + nt * 256;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      uint8_t* tile = smem + (c & 3) * WG_BOX;
      if (c == 4) {                                // the four staging tiles are about to be reused
        if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      uint32_t r[32];
      if (n_iters > 0) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      uint8_t* trow = tile + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        *reinterpret_cast<uint4*>(trow + ((g ^ (row & 7)) << 4)) = make_uint4(r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
      fence_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(&mapC), "r"(smem_u32(tile)), "r"(col0 + c * 32), "r"(m0), "r"((int)blockIdx.y) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
}

// mode-2 problems that fit the pair kernel: whole 256 x 256 tiles, 16-byte aligned partial rows
int launch_wgrad2(const pka_tc_desc* d, cudaStream_t st, bool* handled) {
  *handled = false;
  if (const char* e = getenv("PKA_TC_WGRAD2")) { if (atoi(e) == 0) return PKA_OK; }
  if (d->mode != 2 || d->M % 256 != 0 || d->N % 256 != 0 || d->c_dtype != PKA_F32 || d->ldc % 4 != 0 || !aligned16(d->C)) return PKA_OK;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
    PKA_REQUIRE(e == cudaSuccess, PKA_ELAUNCH, "gemm_tc wgrad2: cannot opt in to shared memory: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  WgradParams p;
  p.Bt = d->Bt; p.T = d->T; p.M = d->M; p.N = d->N; p.nseg = d->nseg; p.ldc = d->ldc;
  for (int i = 0; i < PKA_MAX_CTX; ++i) p.shift[i] = d->shift[i];
  p.tk_per_utt = (d->T + WG_KF - 1) / WG_KF;
  p.n_units = d->Bt * p.tk_per_utt;
  p.units_per_split = (p.n_units + d->splits - 1) / d->splits;
  p.n_tiles_n = d->N / 256;
  p.C = (float*)d->C;
  CUtensorMap mapA, mapB, mapC;
  int rc = make_map(&mapC, d->C, (uint64_t)d->ldc, d->M, d->splits, (uint64_t)d->ldc * 4, (uint64_t)d->M * d->ldc * 4, 128, 1, "gemm_tc dW partials", true);
  if (rc) return rc;
  rc = make_map(&mapA, d->A, d->M, d->T, d->Bt, (uint64_t)d->lda * 2, (uint64_t)d->T * d->lda * 2, WG_KF, 1, "gemm_tc dZ (pair)");
  if (rc) return rc;
  rc = make_map(&mapB, d->B, d->N, d->T, d->Bt, (uint64_t)d->ldb * 2, (uint64_t)d->T * d->ldb * 2, WG_KF, 1, "gemm_tc X (pair)");
  if (rc) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (d->M / 256) * p.n_tiles_n * d->nseg, d->splits, 1);
  cfg.blockDim = dim3(WG_THREADS, 1, 1);
  cfg.dynamicSmemBytes = WG_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_wgrad2_kernel, mapA, mapB, mapC, p);
  PKA_REQUIRE(e == cudaSuccess, PKA_ELAUNCH, "gemm_tc wgrad2 launch failed: %s", cudaGetErrorString(e));
  *handled = true;
  return check_launch("gemm_tc wgrad2");
}

}  // namespace pka
