// "Rows" GEMM of the bf16 tensor-core path, second generation: A-stationary, B-multicast.
//
//   C[b,t,:] = epi( sum_seg A[b, t+shift[seg], :] . W[:, seg*b_seg_col : +K]^T )          (pka_gemm_tc mode 0, a_seg_col == 0)
//
// The first-generation kernel (gemm_tc.cu) streams an A box and a B box per K step; at the TDNN shape
// ([15 968, 3x256] x [768, 256]) that moves 98 MB from L2 to shared memory for 16 MB of operands -- the activations are
// fetched once per splice context and per 128-column output tile, the weights once per CTA -- and the kernel is bound
// by L2 bandwidth at a third of the tensor peak.  Here:
//   * A (activations): one CTA owns a 128-frame tile for ALL output columns and keeps it resident in shared memory,
//     loaded ONCE with an 8-frame halo on each side (TMA zero-fills beyond the utterance = ConcatLayer's padding).  The
//     frame shift of a splice context is a row offset of the UMMA descriptor start address inside the 128B-swizzled
//     tile (matrix base offset = row & 7), so the n_ctx shifted views cost no traffic at all.
//   * B (weights): streamed in [128 x 64] stages through a deep mbarrier ring; the CTAs of a thread-block cluster
//     (different frame tiles, same weights) each fetch 1/CL of every stage and TMA-multicast it to all of them.
//   * accumulators: one 128-column TMEM buffer per 128-column output tile (up to 4); the epilogue of tile n
//     (tcgen05.ld -> bias / ReLU / Philox dropout -> 16-byte stores) overlaps the MMAs of tile n+1.
#include "tc_common.cuh"
#include <stdlib.h>

namespace pka {

constexpr int R2_BM = 128, R2_BN = 128, R2_BK = 64, R2_HALO = 8;
constexpr int R2_B_BYTES = R2_BN * R2_BK * 2;               // 16 KB per stage
constexpr int R2_MAX_STAGES = 8, R2_MAX_NT = 4, R2_MAX_KB = 8;
constexpr int R2_THREADS = 192;

struct Rows2Params {
  int Bt, T, N, K, nseg;
  int kb_per_seg, tiles_per_utt, n_tiles_m;      // n_tiles_m = Bt * tiles_per_utt (grid.x may be padded beyond it)
  int b_seg_col;
  int shift[PKA_MAX_CTX];
  int halo, a_rows, a_blk_bytes;                 // rows per resident A block (128 + 2*halo), bytes per 64-column block
  int stages, nt;                                // B ring depth, number of 128-column output tiles
  int use_base_offset;
  void* C; int ldc, c_dtype;
  const float* bias; int relu;
  pka_dropout drop;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

__global__ void __launch_bounds__(R2_THREADS, 1)
gemm_tc_rows2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const Rows2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int a_bytes = p.kb_per_seg * p.a_blk_bytes;
  uint8_t* sA = smem;
  uint8_t* sB = smem + a_bytes;
  uint64_t* bars = (uint64_t*)(sB + p.stages * R2_B_BYTES);
  // bars: [0,8) a_full per K block, [8,16) b_full, [16,24) b_empty, [24,28) acc_full per output tile; then TMEM slot, bias
  uint64_t* a_full = bars;
  uint64_t* b_full = bars + R2_MAX_KB;
  uint64_t* b_empty = b_full + R2_MAX_STAGES;
  uint64_t* acc_full = b_empty + R2_MAX_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + R2_MAX_NT);
  float* sbias = (float*)(bars + 32);              // 512 floats
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank(), csize = cluster_nctarank();
  const uint16_t cmask = (uint16_t)((1u << csize) - 1u);

  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const bool tile_ok = (int)blockIdx.x < p.n_tiles_m;
  const int b0 = tile_ok ? blockIdx.x / p.tiles_per_utt : p.Bt;                 // padded CTAs read (zero-filled) utterance Bt
  const int t0 = tile_ok ? (blockIdx.x % p.tiles_per_utt) * R2_BM : 0;
  const int nbase = blockIdx.y * (R2_MAX_NT * R2_BN);
  const int n_k = p.nseg * p.kb_per_seg;           // B stages per output tile
  const int total = p.nt * n_k;

  if (threadIdx.x == 0) {
    for (int i = 0; i < R2_MAX_KB; ++i) mbar_init(smem_u32(&a_full[i]), 1);
    for (int s = 0; s < R2_MAX_STAGES; ++s) { mbar_init(smem_u32(&b_full[s]), 1); mbar_init(smem_u32(&b_empty[s]), csize); }
    for (int i = 0; i < R2_MAX_NT; ++i) mbar_init(smem_u32(&acc_full[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();               // every CTA's barriers exist before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {                               // ===== TMA producer
      tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapB);
      for (int kb = 0; kb < p.kb_per_seg; ++kb) {  // resident activations: [a_rows][64] per K block, halo rows included
        mbar_expect_tx(smem_u32(&a_full[kb]), (uint32_t)p.a_blk_bytes);
        tma_load_3d(smem_u32(sA + kb * p.a_blk_bytes), &mapA, smem_u32(&a_full[kb]), kb * R2_BK, t0 - p.halo, b0);
      }
      const int rows_per_cta = R2_BN / (int)csize;
      for (int i = 0; i < total; ++i) {
        const int s = i % p.stages, round = i / p.stages;
        mbar_wait(smem_u32(&b_empty[s]), (round & 1) ^ 1);      // all CTAs of the cluster have released this stage
        const uint32_t full = smem_u32(&b_full[s]);
        mbar_expect_tx(full, R2_B_BYTES);
        const int nt = i / n_k, kk = i % n_k;
        const int seg = kk / p.kb_per_seg, kb = kk % p.kb_per_seg;
        const int col = seg * p.b_seg_col + kb * R2_BK;
        const int row = nbase + nt * R2_BN + (int)rank * rows_per_cta;
        const uint32_t dst = smem_u32(sB + s * R2_B_BYTES + (int)rank * rows_per_cta * 128);
        if (csize > 1) tma_load_3d_mc(dst, &mapB, full, col, row, 0, cmask);
        else tma_load_3d(dst, &mapB, full, col, row, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                               // ===== MMA issuer
      const uint32_t idesc = make_idesc(R2_BM, R2_BN);
      for (int i = 0; i < total; ++i) {
        const int s = i % p.stages, round = i / p.stages;
        const int nt = i / n_k, kk = i % n_k;
        const int seg = kk / p.kb_per_seg, kb = kk % p.kb_per_seg;
        if (nt == 0 && seg == 0) mbar_wait(smem_u32(&a_full[kb]), 0);
        mbar_wait(smem_u32(&b_full[s]), round & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + kb * p.a_blk_bytes) + (uint32_t)(p.halo + p.shift[seg]) * 128u;
        uint64_t da = make_sdesc(a_addr);
        if (p.use_base_offset) da |= (uint64_t)((a_addr >> 7) & 7u) << 49;
        const uint64_t db = make_sdesc(smem_u32(sB + s * R2_B_BYTES));
#pragma unroll
        for (int k = 0; k < R2_BK / 16; ++k)
          umma_f16(tmem_base + (uint32_t)(nt * R2_BN), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kk | k) ? 1u : 0u);
        if (csize > 1) umma_commit_mc(smem_u32(&b_empty[s]), cmask);
        else umma_commit(smem_u32(&b_empty[s]));
        if (kk == n_k - 1) umma_commit(smem_u32(&acc_full[nt]));
      }
    }
  } else {                                         // ===== epilogue warps 2..5 -> TMEM lane quarter (warp % 4)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    for (int e = threadIdx.x - 64; e < p.nt * R2_BN; e += 128) {
      const int n = nbase + e;
      sbias[e] = (p.bias && n < p.N) ? p.bias[n] : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const DropCtx dc = make_drop(p.drop);
    const float relu_floor = p.relu ? 0.f : -3.4e38f;
    const int t = t0 + row;
    const bool valid = tile_ok && t < p.T;
    const long long m = (long long)b0 * p.T + t;
    for (int nt = 0; nt < p.nt; ++nt) {
      mbar_wait(smem_u32(&acc_full[nt]), 0);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < R2_BN / 32; ++c) {
        const int nb = nbase + nt * R2_BN + c * 32;
        if (nb >= p.N) break;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(nt * R2_BN + c * 32), r);
        const bool full = nb + 32 <= p.N;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bv = *reinterpret_cast<const float4*>(&sbias[nt * R2_BN + c * 32 + j]);
          v[j] = fmaxf(__uint_as_float(r[j]) + bv.x, relu_floor);
          v[j + 1] = fmaxf(__uint_as_float(r[j + 1]) + bv.y, relu_floor);
          v[j + 2] = fmaxf(__uint_as_float(r[j + 2]) + bv.z, relu_floor);
          v[j + 3] = fmaxf(__uint_as_float(r[j + 3]) + bv.w, relu_floor);
        }
        if (dc.p > 0.f) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (nb + j < p.N) {
              const float4 mul = dropout_mul4(dc, ((unsigned long long)m * p.N + nb + j) >> 2);
              v[j] *= mul.x; v[j + 1] *= mul.y; v[j + 2] *= mul.z; v[j + 3] *= mul.w;
            }
          }
        }
        if (valid) {
          if (p.c_dtype == PKA_BF16) {
            __nv_bfloat16* dst = (__nv_bfloat16*)p.C + m * p.ldc + nb;
            if (full && (p.ldc & 7) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 pk;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]), h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                pk.x = *(uint32_t*)&h0; pk.y = *(uint32_t*)&h1; pk.z = *(uint32_t*)&h2; pk.w = *(uint32_t*)&h3;
                *reinterpret_cast<uint4*>(dst + j) = pk;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (nb + j < p.N) dst[j] = __float2bfloat16_rn(v[j]);
            }
          } else {
            float* dst = (float*)p.C + m * p.ldc + nb;
            if (full && (p.ldc & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (nb + j < p.N) dst[j] = v[j];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();               // nobody leaves while a peer may still multicast into / arrive on its smem
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// Can this mode-0 problem run on the A-stationary kernel?  On success fills the launch geometry.
struct Rows2Plan { int halo, a_rows, a_blk_bytes, stages, nt, smem, cluster, grid_x, grid_y; };
static bool plan_rows2(const pka_tc_desc* d, Rows2Plan* pl) {
  if (d->mode != 0 || d->Ct) return false;
  if (d->nseg > 1 && d->a_seg_col != 0) return false;
  const int kb = (d->K + R2_BK - 1) / R2_BK;
  if (kb > R2_MAX_KB) return false;
  int halo = 0;
  for (int i = 0; i < d->nseg; ++i) {
    const int s = d->shift[i] < 0 ? -d->shift[i] : d->shift[i];
    if (s > R2_HALO) return false;
    if (s) halo = R2_HALO;
  }
  pl->halo = halo;
  pl->a_rows = R2_BM + 2 * halo;
  pl->a_blk_bytes = pl->a_rows * 128;
  const int a_bytes = kb * pl->a_blk_bytes;
  const int budget = 225 * 1024 - a_bytes - 256 - 2048 - 64;
  int stages = budget / R2_B_BYTES;
  if (stages > R2_MAX_STAGES) stages = R2_MAX_STAGES;
  if (stages < 2) return false;
  const int n_tiles_n = (d->N + R2_BN - 1) / R2_BN;
  pl->nt = n_tiles_n < R2_MAX_NT ? n_tiles_n : R2_MAX_NT;
  pl->grid_y = (n_tiles_n + R2_MAX_NT - 1) / R2_MAX_NT;
  if (pl->grid_y > 1 && n_tiles_n % R2_MAX_NT != 0) return false;      // keep every grid row on the same tile count
  const int total = pl->nt * d->nseg * kb;
  if (stages > total) stages = total;
  pl->stages = stages;
  pl->smem = a_bytes + stages * R2_B_BYTES + 256 + 2048;
  if (pl->smem < 116 * 1024) pl->smem = 116 * 1024;                      // one CTA per SM: it owns all 512 TMEM columns
  const int tiles_m = d->Bt * ((d->T + R2_BM - 1) / R2_BM);
  int cl = tiles_m >= 64 ? 4 : (tiles_m >= 8 ? 2 : 1);
  if (const char* e = getenv("PKA_TC_CLUSTER")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) cl = v; }
  pl->cluster = cl;
  pl->grid_x = (tiles_m + cl - 1) / cl * cl;
  return true;
}

int launch_rows2(const pka_tc_desc* d, cudaStream_t st, bool* handled) {
  *handled = false;
  if (const char* e = getenv("PKA_TC_ROWS2")) { if (atoi(e) == 0) return PKA_OK; }
  Rows2Plan pl;
  if (!plan_rows2(d, &pl)) return PKA_OK;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_rows2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    PKA_REQUIRE(e == cudaSuccess, PKA_ELAUNCH, "gemm_tc rows2: cannot opt in to shared memory: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  Rows2Params p;
  p.Bt = d->Bt; p.T = d->T; p.N = d->N; p.K = d->K; p.nseg = d->nseg;
  p.kb_per_seg = (d->K + R2_BK - 1) / R2_BK;
  p.tiles_per_utt = (d->T + R2_BM - 1) / R2_BM;
  p.n_tiles_m = d->Bt * p.tiles_per_utt;
  p.b_seg_col = d->b_seg_col;
  for (int i = 0; i < PKA_MAX_CTX; ++i) p.shift[i] = d->shift[i];
  p.halo = pl.halo; p.a_rows = pl.a_rows; p.a_blk_bytes = pl.a_blk_bytes;
  p.stages = pl.stages; p.nt = pl.nt;
  { const char* e = getenv("PKA_TC_BASEOFF"); p.use_base_offset = e ? atoi(e) : 1; }
  p.C = d->C; p.ldc = d->ldc; p.c_dtype = d->c_dtype; p.bias = d->bias; p.relu = d->relu; p.drop = d->drop;
  CUtensorMap mapA, mapB;
  int rc = make_map(&mapA, d->A, (uint64_t)d->K, d->T, d->Bt, (uint64_t)d->lda * 2, (uint64_t)d->T * d->lda * 2, pl.a_rows, 1, "gemm_tc A (resident)");
  if (rc) return rc;
  const uint64_t b_cols = (uint64_t)d->b_seg_col * (d->nseg - 1) + d->K;
  rc = make_map(&mapB, d->B, b_cols, d->N, 1, (uint64_t)d->ldb * 2, (uint64_t)d->N * d->ldb * 2, R2_BN / pl.cluster, 1, "gemm_tc B (multicast)");
  if (rc) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.grid_x, pl.grid_y, 1);
  cfg.blockDim = dim3(R2_THREADS, 1, 1);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pl.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_rows2_kernel, mapA, mapB, p);
  PKA_REQUIRE(e == cudaSuccess, PKA_ELAUNCH, "gemm_tc rows2 launch failed: %s", cudaGetErrorString(e));
  *handled = true;
  return check_launch("gemm_tc rows2");
}

}  // namespace pka
