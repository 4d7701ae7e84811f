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
//     frame shift of a splice context is a row offset (128 B per frame) of the UMMA descriptor start address inside the
//     128B-swizzled tile, so the n_ctx shifted views cost no traffic at all.  (Measured on B200: the swizzle XOR is
//     taken from the absolute shared-memory address bits, so a start address that is not 1 KB aligned needs NO
//     "matrix base offset" in the descriptor -- setting it to (addr >> 7) & 7 gives wrong products.)
//   * B (weights): streamed in [128 x 64] stages through a deep mbarrier ring; the CTAs of a thread-block cluster
//     (different frame tiles, same weights) each fetch 1/CL of every stage and TMA-multicast it to all of them.
//   * accumulators: one 128-column TMEM buffer per 128-column output tile (up to 4); the epilogue of tile n
//     (tcgen05.ld -> bias / ReLU / Philox dropout -> 16-byte stores) overlaps the MMAs of tile n+1.
#include "tc_common.cuh"
#include <stdlib.h>

namespace pka {

constexpr int R2_BM = 128, R2_BN = 128, R2_BK = 64, R2_HALO = 8;
constexpr int R2_BOX_BYTES = R2_BN * R2_BK * 2;             // one [128 x 64] weight box = 16 KB
constexpr int R2_STAGE_BYTES = 2 * R2_BOX_BYTES;            // a stage holds two consecutive K blocks (K = 128)
constexpr int R2_C_BYTES = R2_BM * 64 * 2;                  // output staging tile [128 rows][64 bf16] for the TMA store
constexpr int R2_MAX_STAGES = 6, R2_MAX_NT = 4, R2_MAX_KB = 8;
constexpr int R2_THREADS = 352;                             // warp 0 / warp 10 TMA (even / odd stages), warp 1 MMA, warps 2-9 epilogue

struct Rows2Params {
  int Bt, T, N, K, nseg;
  int kb_per_seg, tiles_per_utt, n_tiles_m;      // n_tiles_m = Bt * tiles_per_utt (grid.x may be padded beyond it)
  int b_seg_col;
  int shift[PKA_MAX_CTX];
  int halo, a_rows, a_blk_bytes;                 // rows per resident A block (128 + 2*halo), bytes per 64-column block
  int stages, nt;                                // B ring depth, number of output-column chunks (one TMEM accumulator each)
  int chunk, stage_bytes;                        // columns per chunk (128; pair mode: up to 256), bytes of one B stage per CTA
  int use_base_offset, dbg;                      // diagnostics (env PKA_TC_BASEOFF / PKA_TC_DBG: 1 skip MMAs, 4 skip stores, 8 skip epilogue, 16 stamps)
  int tma_store;                                 // bf16 output through shared memory + cp.async.bulk.tensor stores
  void* C; int ldc, c_dtype;
  const float* bias; int relu;
  pka_dropout drop;
  const __nv_bfloat16* addend; int ld_add;       // optional: C = epi(acc) + addend[m, n]  (residual-gradient branch)
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_relaxed() {     // execution barrier only (no memory fence: no MEMBAR.ALL.GPU)
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
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
// CTA-pair (cta_group::2) variants: the load signals the LEADER CTA's mbarrier (peer bit cleared), the MMA spans both
// SMs (M = 256: each CTA's tensor core works on its own 128 rows, B rows are split between the two shared memories)
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// diagnostics (PKA_TC_DBG & 16): per-CTA globaltimer stamps, read back with pka_debug_rows2_stamps()
__device__ unsigned long long g_rows2_ts[256 * 16];
__device__ __forceinline__ void stamp(int dbg, int slot) {
  if (dbg & 16) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (blockIdx.x < 256 && blockIdx.y == 0) g_rows2_ts[blockIdx.x * 16 + slot] = t;
  }
}

template <bool PAIR>
__global__ void __launch_bounds__(R2_THREADS, 1)
gemm_tc_rows2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                     const __grid_constant__ CUtensorMap mapC, const Rows2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int a_bytes = p.kb_per_seg * p.a_blk_bytes;
  uint8_t* sA = smem;
  uint8_t* sB = smem + a_bytes;
  uint8_t* sC = sB + p.stages * p.stage_bytes;
  uint64_t* bars = (uint64_t*)(sC + 2 * R2_C_BYTES);
  // bars: [0,8) a_full per K block, [8,14) b_full, [14,20) b_empty, [20,24) acc_full per output tile; then TMEM slot, bias
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
  if (threadIdx.x == 0) stamp(p.dbg, 0);

  const bool tile_ok = (int)blockIdx.x < p.n_tiles_m;
  const int b0 = tile_ok ? blockIdx.x / p.tiles_per_utt : p.Bt;                 // padded CTAs read (zero-filled) utterance Bt
  const int t0 = tile_ok ? (blockIdx.x % p.tiles_per_utt) * R2_BM : 0;
  const int nbase = blockIdx.y * (p.nt * p.chunk);
  const int box_bytes = p.stage_bytes >> 1;
  const bool leader = !PAIR || rank == 0;
  const int kp_per_seg = (p.kb_per_seg + 1) >> 1;  // stages per splice context (two K blocks each, the last may hold one)

  if (threadIdx.x == 0) {
    for (int i = 0; i < R2_MAX_KB; ++i) mbar_init(smem_u32(&a_full[i]), 1);
    for (int s = 0; s < R2_MAX_STAGES; ++s) { mbar_init(smem_u32(&b_full[s]), 1); mbar_init(smem_u32(&b_empty[s]), PAIR ? 1u : csize); }
    for (int i = 0; i < R2_MAX_NT; ++i) mbar_init(smem_u32(&acc_full[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc<512>(smem_u32(tmem_slot));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_relaxed();           // every CTA's barriers exist (fence.mbarrier_init above) before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, cluster hand-shake) may overlap the
  // tail of the previous kernel in the stream; nothing below touches global memory before that kernel has completed.
  pdl_wait();
  if (threadIdx.x == 0) stamp(p.dbg, 1);

  if (warp == 0 || warp == 10) {
    const int pid = warp == 0 ? 0 : 1;             // two producer threads on different schedulers: even / odd stages
    if (lane == 0) {                               // ===== TMA producer (no divisions in the loop: it paces the whole CTA)
      tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapB);
      const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
      const int rows_per_cta = PAIR ? (p.chunk >> 1) : R2_BN / (int)csize;
      const uint32_t slice_off = PAIR ? 0u : rank * (uint32_t)rows_per_cta * 128u;
      int a_next = 0;
      auto issue_a = [&]() {                       // resident activations: [a_rows][64] per K block, halo rows included
        const uint32_t bar = smem_u32(&a_full[a_next]);
        if (PAIR) {                                // both CTAs' tiles complete on the leader's barrier
          if (leader) mbar_expect_tx(bar, 2u * (uint32_t)p.a_blk_bytes);
          tma_load_3d_pair(sA_u + a_next * p.a_blk_bytes, &mapA, bar, a_next * R2_BK, t0 - p.halo, b0);
        } else {
          mbar_expect_tx(bar, (uint32_t)p.a_blk_bytes);
          tma_load_3d(sA_u + a_next * p.a_blk_bytes, &mapA, bar, a_next * R2_BK, t0 - p.halo, b0);
        }
        ++a_next;
      };
      if (pid == 0) {
        issue_a();
        if (p.kb_per_seg > 1) issue_a();
      } else {
        a_next = p.kb_per_seg;                     // the activations are producer 0's job
      }
      int s = 0, istage = 0;
      uint32_t ph = 0;
      for (int nt = 0; nt < p.nt; ++nt) {
        const int row = nbase + nt * p.chunk + (int)rank * rows_per_cta;
        for (int seg = 0; seg < p.nseg; ++seg) {
          int col = seg * p.b_seg_col, kb_left = p.kb_per_seg;
          for (int kp = 0; kp < kp_per_seg; ++kp, col += 2 * R2_BK, kb_left -= 2, ++istage) {
            if ((istage & 1) != pid) { if (++s == p.stages) { s = 0; ph ^= 1u; } continue; }
            mbar_wait(smem_u32(&b_empty[s]), ph ^ 1u);          // every consumer of this stage has released it
            const uint32_t full = smem_u32(&b_full[s]);
            const uint32_t dst = sB_u + (uint32_t)s * p.stage_bytes + slice_off;
            const int nbx = kb_left >= 2 ? 2 : 1;
            if (PAIR) {
              if (leader) mbar_expect_tx(full, 2u * (uint32_t)(nbx * box_bytes));
              tma_load_3d_pair(dst, &mapB, full, col, row, 0);
              if (nbx == 2) tma_load_3d_pair(dst + box_bytes, &mapB, full, col + R2_BK, row, 0);
            } else {
              mbar_expect_tx(full, (uint32_t)(nbx * box_bytes));
              if (csize > 1) {
                tma_load_3d_mc(dst, &mapB, full, col, row, 0, cmask);
                if (nbx == 2) tma_load_3d_mc(dst + box_bytes, &mapB, full, col + R2_BK, row, 0, cmask);
              } else {
                tma_load_3d(dst, &mapB, full, col, row, 0);
                if (nbx == 2) tma_load_3d(dst + box_bytes, &mapB, full, col + R2_BK, row, 0);
              }
            }
            // producer 0 owns the activations but only every other weight stage: after its turn at stage i the K
            // blocks of stages i+1 and i+2 must be on their way (four blocks per turn; two were not enough when a
            // problem has fewer weight stages than K-block pairs, e.g. K = 512 -> N = 128: blocks 6 and 7 never came)
#pragma unroll 1
            for (int r = 0; r < 4 && a_next < p.kb_per_seg; ++r) issue_a();
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
      while (a_next < p.kb_per_seg) issue_a();     // (pid 1 starts at kb_per_seg: no-op there)
      if (pid == 0) stamp(p.dbg, 4);
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {                     // ===== MMA issuer (pair mode: the leader CTA issues for both SMs)
      const uint32_t idesc = PAIR ? make_idesc(2 * R2_BM, p.chunk) : make_idesc(R2_BM, R2_BN);
      const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
      int s = 0;
      uint32_t ph = 0;
      for (int nt = 0; nt < p.nt; ++nt) {
        const uint32_t tmem_d = tmem_base + (uint32_t)(nt * p.chunk);
        for (int seg = 0; seg < p.nseg; ++seg) {
          const uint32_t a_off = (uint32_t)(p.halo + p.shift[seg]) * 128u;
          int kb = 0;
          for (int kp = 0; kp < kp_per_seg; ++kp) {
            const int nb = min(2, p.kb_per_seg - kb);
            if (nt == 0 && seg == 0) {
              for (int j = 0; j < nb; ++j) mbar_wait(smem_u32(&a_full[kb + j]), 0);
              if (kp == 0) stamp(p.dbg, 5);
            }
            mbar_wait(smem_u32(&b_full[s]), ph);
            if (nt == 0 && seg == 0 && kp == 0) stamp(p.dbg, 6);
            tc_fence_after();
            for (int j = 0; j < nb; ++j, ++kb) {
              const uint32_t a_addr = sA_u + (uint32_t)(kb * p.a_blk_bytes) + a_off;
              uint64_t da = make_sdesc(a_addr);
              if (p.use_base_offset) da |= (uint64_t)((a_addr >> 7) & 7u) << 49;
              const uint64_t db = make_sdesc(sB_u + (uint32_t)s * p.stage_bytes + (uint32_t)(j * box_bytes));
              if (!(p.dbg & 1)) {
#pragma unroll
                for (int k = 0; k < R2_BK / 16; ++k) {
                  if (PAIR) umma_f16_pair(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (seg | kp | j | k) ? 1u : 0u);
                  else umma_f16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (seg | kp | j | k) ? 1u : 0u);
                }
              }
            }
            if (PAIR) umma_commit_pair(smem_u32(&b_empty[s]));
            else if (csize > 1) umma_commit_mc(smem_u32(&b_empty[s]), cmask);
            else umma_commit(smem_u32(&b_empty[s]));
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
        if (PAIR) umma_commit_pair(smem_u32(&acc_full[nt]));
        else umma_commit(smem_u32(&acc_full[nt]));
        stamp(p.dbg, 7 + (nt & 1));
      }
    }
  } else {                                         // ===== epilogue warps 2..9: TMEM lane quarter = warp % 4, column group = (warp-2)/4
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;               // group g owns the 64-column slices g, g+2, ... of every accumulator
    const int gtid = (int)threadIdx.x - 64 - grp * 128;
    const int row = q * 32 + lane;
    for (int e = threadIdx.x - 64; e < p.nt * p.chunk; e += 256) {
      const int n = nbase + e;
      sbias[e] = (p.bias && n < p.N) ? p.bias[n] : 0.f;
    }
    asm volatile("bar.sync 3, 256;" ::: "memory");
    const DropCtx dc = make_drop(p.drop);
    const float relu_floor = p.relu ? 0.f : -3.4e38f;
    const int t = t0 + row;
    const bool valid = tile_ok && t < p.T && !(p.dbg & 4);
    const long long m = (long long)b0 * p.T + t;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint8_t* sCg = sC + grp * R2_C_BYTES;
    const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
    // Once the LAST accumulator is complete every MMA has read the weight ring and nothing writes it any more, so the
    // ring itself serves as staging: each 64-column slice of the last chunk gets its own 16 KB tile and its bulk store
    // never has to wait for an earlier one (the two per-group tiles of sC serve the chunks that overlap the MMAs).
    const bool ring_ok = p.stages * p.stage_bytes >= (p.chunk >> 6) * R2_C_BYTES;
    bool store_pending = false;
    // Dropout keep bits depend only on (row, column), not on the accumulators: generate them NOW, while the MMAs run,
    // so that the Philox rounds are off the exposed epilogue (1 bit per element, <= 256 columns per thread).
    uint32_t keep[8];
    if (dc.p > 0.f) {
      int w = 0;
      for (int nt = 0; nt < p.nt; ++nt)
        for (int hf = grp; hf < (p.chunk >> 6); hf += 2) {
          const int nb0 = nbase + nt * p.chunk + hf * 64;
#pragma unroll
          for (int c = 0; c < 2; ++c, ++w) {
            uint32_t bits = 0u;
            if ((p.N & 7) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const int nb = nb0 + c * 32 + j;
                if (nb < p.N) bits |= dropout_bits8(dc, ((unsigned long long)m * p.N + nb) >> 3) << j;
              }
            } else {
#pragma unroll 4
              for (int j = 0; j < 32; ++j) {
                const int nb = nb0 + c * 32 + j;
                if (nb < p.N && dropout_keep(dc, (unsigned long long)m * p.N + nb)) bits |= 1u << j;
              }
            }
#pragma unroll
            for (int z = 0; z < 8; ++z) if (w == z) keep[z] = bits;
          }
        }
    }
    int kw = 0;
    for (int nt = 0; nt < p.nt; ++nt) {
      mbar_wait(smem_u32(&acc_full[nt]), 0);
      if (threadIdx.x == 64) stamp(p.dbg, 10 + (nt & 1));
      if (threadIdx.x == 64 && nt == p.nt - 1) stamp(p.dbg, 12);
      tc_fence_after();
      if (p.dbg & 8) continue;
#pragma unroll 1
      for (int hf = grp; hf < (p.chunk >> 6); hf += 2) { // 64 output columns at a time
        const int nb0 = nbase + nt * p.chunk + hf * 64;
        if (nb0 >= p.N) break;
        const bool in_ring = ring_ok && nt == p.nt - 1;
        uint8_t* tile = in_ring ? sB + hf * R2_C_BYTES : sCg;
        uint8_t* crow = tile + row_off;
        if (p.tma_store && !in_ring) {
          if (store_pending) {                     // the previous bulk store must have read the staging tile
            if (gtid == 0) tma_store_wait_read();
            if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
          }
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int nb = nb0 + c * 32;
          uint32_t r[32];
          tmem_ld32(tmem_base + lane_addr + (uint32_t)(nt * p.chunk + hf * 64 + c * 32), r);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(&sbias[nt * p.chunk + hf * 64 + c * 32 + j]);
            v[j] = fmaxf(__uint_as_float(r[j]) + bv.x, relu_floor);
            v[j + 1] = fmaxf(__uint_as_float(r[j + 1]) + bv.y, relu_floor);
            v[j + 2] = fmaxf(__uint_as_float(r[j + 2]) + bv.z, relu_floor);
            v[j + 3] = fmaxf(__uint_as_float(r[j + 3]) + bv.w, relu_floor);
          }
          if (dc.p > 0.f) {
            uint32_t bits = keep[0];
#pragma unroll
            for (int z = 1; z < 8; ++z) if (kw == z) bits = keep[z];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = ((bits >> j) & 1u) ? v[j] * dc.scale : 0.f;
          }
          ++kw;
          if (p.addend && valid && nb < p.N) {     // + residual-gradient branch (bf16 row segment of this thread's row)
            const __nv_bfloat16* ar = p.addend + m * p.ld_add + nb;
            if (nb + 32 <= p.N && (p.ld_add & 7) == 0) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint4 a4 = *reinterpret_cast<const uint4*>(ar + g * 8);
                const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a4);
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(ah[k]); v[g * 8 + 2 * k] += f.x; v[g * 8 + 2 * k + 1] += f.y; }
              }
            } else {                               // (fully unrolled: a rolled loop would index v[] dynamically and
#pragma unroll                                   //  push the whole accumulator row into local memory)
              for (int j = 0; j < 32; ++j) if (nb + j < p.N) v[j] += __bfloat162float(ar[j]);
            }
          }
          if (p.tma_store) {                       // bf16 row segment -> swizzled staging tile (conflict-free 16 B stores)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 pk;
              __nv_bfloat162 h0 = __floats2bfloat162_rn(v[g * 8], v[g * 8 + 1]), h1 = __floats2bfloat162_rn(v[g * 8 + 2], v[g * 8 + 3]);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(v[g * 8 + 4], v[g * 8 + 5]), h3 = __floats2bfloat162_rn(v[g * 8 + 6], v[g * 8 + 7]);
              pk.x = *(uint32_t*)&h0; pk.y = *(uint32_t*)&h1; pk.z = *(uint32_t*)&h2; pk.w = *(uint32_t*)&h3;
              *reinterpret_cast<uint4*>(crow + (((c * 4 + g) ^ (row & 7)) << 4)) = pk;
            }
          } else if (valid && nb < p.N) {
            const bool full = nb + 32 <= p.N;
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
        if (p.tma_store) {
          fence_async_smem();
          if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
          if (gtid == 0 && tile_ok && !(p.dbg & 4)) tma_store_3d(&mapC, smem_u32(tile), nb0, t0, b0);   // rows >= T are clipped
          if (!in_ring) store_pending = true;
        }
      }
    }
    if (p.tma_store && gtid == 0) tma_store_wait_read();
    if (threadIdx.x == 64) stamp(p.dbg, 14);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) stamp(p.dbg, 15);
  if (csize > 1) cluster_sync_relaxed();           // nobody leaves while a peer may still multicast into / arrive on its smem
  if (warp == 2) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    else tmem_dealloc<512>(tmem_base);
  }
}

// Can this mode-0 problem run on the A-stationary kernel?  On success fills the launch geometry.
struct Rows2Plan { int halo, a_rows, a_blk_bytes, stages, nt, chunk, stage_bytes, pair, smem, cluster, grid_x, grid_y; };
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
  const int tiles_m = d->Bt * ((d->T + R2_BM - 1) / R2_BM);
  // CTA-pair mode (cta_group::2, M = 256): each CTA stages only HALF of every weight tile, which halves the dominant
  // shared-memory ingress.  Needs equal column chunks of <= 256 that are multiples of 64, all in one CTA's 512 TMEM columns.
  int pair = 0, nt = 0, chunk = R2_BN, pair_gy = 1;
  const int budget0 = 225 * 1024 - a_bytes - 2 * R2_C_BYTES - 256 - 2048 - 64;
  if (tiles_m >= 2) {
    // Column chunk candidates, widest first (16 KB TMA boxes); a CTA covers up to 512 columns (its TMEM), wider outputs
    // go over grid.y.  The weight ring must hold >= 3 stages next to the resident activations (>= 2 as a last resort):
    // with d_model = 512 activations (128 KB) only the 128-column chunks (16 KB stages) fit.
    int cand[3] = {256, 128, (d->N <= 256 && d->N % 64 == 0) ? d->N : 0};
    if (const char* e = getenv("PKA_TC_CHUNK")) { const int c = atoi(e); if (c >= 64 && c <= 256 && c % 64 == 0) { cand[0] = c; cand[1] = 0; cand[2] = 0; } }
    for (int need = 3; need >= 2 && !pair; --need)
      for (int ci = 0; ci < 3 && !pair; ++ci) {
        const int c = cand[ci];
        if (c <= 0 || d->N % c != 0) continue;
        // columns per CTA: the largest multiple of the chunk that divides N and fits the 512 TMEM columns
        // (N = 768, the fused cross-attention k|v projection of three layers: 256 per CTA, three grid rows)
        int per_cta = (d->N < 512 ? d->N : 512) / c * c;
        while (per_cta >= c && d->N % per_cta != 0) per_cta -= c;
        if (per_cta < c) continue;
        if (budget0 / (c * 128) < need) continue;
        pair = 1; chunk = c; nt = per_cta / c; pair_gy = d->N / per_cta;
      }
  }
  if (const char* e = getenv("PKA_TC_PAIR")) pair = pair && atoi(e) != 0;
  pl->pair = pair;
  if (pair) {
    pl->nt = nt; pl->chunk = chunk; pl->grid_y = pair_gy;
    pl->stage_bytes = chunk * 128;                                       // two boxes of [chunk/2 rows][64]
  } else {
    const int n_tiles_n = (d->N + R2_BN - 1) / R2_BN;
    pl->nt = n_tiles_n < R2_MAX_NT ? n_tiles_n : R2_MAX_NT;
    pl->chunk = R2_BN;
    pl->grid_y = (n_tiles_n + R2_MAX_NT - 1) / R2_MAX_NT;
    if (pl->grid_y > 1 && n_tiles_n % R2_MAX_NT != 0) return false;    // keep every grid row on the same tile count
    pl->stage_bytes = R2_STAGE_BYTES;
  }
  int stages = budget0 / pl->stage_bytes;
  if (stages > R2_MAX_STAGES) stages = R2_MAX_STAGES;
  if (stages < 2) return false;
  const int total = pl->nt * d->nseg * ((kb + 1) / 2);
  if (stages > total) stages = total;
  const int ring_min = ((pl->chunk >> 6) * R2_C_BYTES + pl->stage_bytes - 1) / pl->stage_bytes;   // epilogue staging of the last chunk
  if (stages < ring_min && ring_min * pl->stage_bytes <= budget0) stages = ring_min;
  pl->stages = stages;
  pl->smem = a_bytes + stages * pl->stage_bytes + 2 * R2_C_BYTES + 256 + 2048;
  if (pl->smem < 116 * 1024) pl->smem = 116 * 1024;                      // one CTA per SM: it owns all 512 TMEM columns
  int cl = pair ? 2 : (tiles_m >= 64 ? 4 : (tiles_m >= 8 ? 2 : 1));
  if (!pair) if (const char* e = getenv("PKA_TC_CLUSTER")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) cl = v; }
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
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_rows2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_rows2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
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
  p.stages = pl.stages; p.nt = pl.nt; p.chunk = pl.chunk; p.stage_bytes = pl.stage_bytes;
  { const char* e = getenv("PKA_TC_DBG"); p.dbg = e ? atoi(e) : 0; }
  { const char* e = getenv("PKA_TC_BASEOFF"); p.use_base_offset = e ? atoi(e) : 0; }     // diagnostics only, see header
  p.C = d->C; p.ldc = d->ldc; p.c_dtype = d->c_dtype; p.bias = d->bias; p.relu = d->relu; p.drop = d->drop;
  p.addend = (const __nv_bfloat16*)d->addend; p.ld_add = d->ldadd;
  p.tma_store = (d->c_dtype == PKA_BF16 && d->ldc % 8 == 0 && aligned16(d->C)) ? 1 : 0;
  if (const char* e = getenv("PKA_TC_TMASTORE")) p.tma_store = p.tma_store && atoi(e) != 0;
  CUtensorMap mapA, mapB, mapC;
  int rc = make_map(&mapA, d->A, (uint64_t)d->K, d->T, d->Bt, (uint64_t)d->lda * 2, (uint64_t)d->T * d->lda * 2, pl.a_rows, 1, "gemm_tc A (resident)");
  if (rc) return rc;
  const uint64_t b_cols = (uint64_t)d->b_seg_col * (d->nseg - 1) + d->K;
  rc = make_map(&mapB, d->B, b_cols, d->N, 1, (uint64_t)d->ldb * 2, (uint64_t)d->N * d->ldb * 2, pl.pair ? pl.chunk / 2 : R2_BN / pl.cluster, 1, "gemm_tc B");
  if (rc) return rc;
  if (p.tma_store) {
    rc = make_map(&mapC, d->C, (uint64_t)d->N, d->T, d->Bt, (uint64_t)d->ldc * 2, (uint64_t)d->T * d->ldc * 2, R2_BM, 1, "gemm_tc C (store)");
    if (rc) return rc;
  } else {
    mapC = mapA;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.grid_x, pl.grid_y, 1);
  cfg.blockDim = dim3(R2_THREADS, 1, 1);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pl.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (pdl_enabled()) cfg.numAttrs = 2;
  cudaError_t e = pl.pair ? cudaLaunchKernelEx(&cfg, gemm_tc_rows2_kernel<true>, mapA, mapB, mapC, p)
                          : cudaLaunchKernelEx(&cfg, gemm_tc_rows2_kernel<false>, mapA, mapB, mapC, p);
  PKA_REQUIRE(e == cudaSuccess, PKA_ELAUNCH, "gemm_tc rows2 launch failed: %s", cudaGetErrorString(e));
  *handled = true;
  return check_launch("gemm_tc rows2");
}

}  // namespace pka

extern "C" int pka_debug_rows2_stamps(unsigned long long* host, int n_bytes) {
  return cudaMemcpyFromSymbol(host, pka::g_rows2_ts, n_bytes) == cudaSuccess ? 0 : 4;
}
