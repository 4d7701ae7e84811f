// Banded / key-padded multi-head attention on the 5th-generation tensor cores (sm_100a), forward pass.
// replaces: ScaledDotProductAttention.forward (T/Modules.py:75-97) + the masks of T/Models.py:27-49 + the head
// replication of T/SubLayers.py:49-59, for bf16 activations with head dim 64 (the only head dim the reference's
// Transformer can build: T/Models.py:250-251 never forwards d_k/d_v to the Decoder).
//
// One CTA = one (utterance, head, 128-query tile); key tiles of 128 stream through a 2-stage TMA ring:
//   warp 4 lane 0   TMA producer: Q once, then K_j / V_j boxes [128 keys][64] (128B swizzle, OOB rows zero-filled)
//   warp 5 lane 0   MMA issuer:   S = Q K_j^T   (tcgen05.mma 128x128x16 x4, both operands K-major)  -> TMEM cols [0,128)
//                                 O_j = P_j V_j (tcgen05.mma 128x64x16  x8, P K-major from smem, V MN-major) -> TMEM cols [128,192)
//   warps 0-3       online softmax, one thread per query row (TMEM lane = row): ONE tcgen05.ld pass brings the 128 scores
//                   of the tile into registers, mask predicate (key padding via warp ballot of the u8 mask, band via
//                   per-row bit range; chunks of 32 keys that are all allowed skip the per-element tests), running
//                   reference maximum / sum in the exp2 domain, P_j written to shared memory as bf16 in the UMMA K-major
//                   128B-swizzle layout.
//   O stays in TENSOR MEMORY for the whole key loop: P_j V_j accumulates into it (round 1 pulled every partial product
//   into registers and rescaled 64 accumulators per tile).  The exponent reference m_ref is only raised -- and O / l
//   rescaled through tcgen05.ld / tcgen05.st -- when a tile's maximum exceeds it by more than 2^8 (lazy rescale: the
//   probabilities then stay below 256, far inside bf16 / fp32 range, and the result is mathematically the same).
// Two CTAs are resident per SM (112 KB smem, 256 TMEM columns each), so one CTA's softmax overlaps the other's MMAs.
// The [B*H, Lq, Lk] probability tensor never exists; rows without an allowed key give 0 output and lse = -inf.
#include "tc_common.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace pka {

constexpr int AT_BM = 128, AT_BN = 128, AT_D = 64, AT_STAGES = 2;
constexpr int AT_Q_BYTES = AT_BM * AT_D * 2;                  // 16 KB
constexpr int AT_KV_BYTES = AT_BN * AT_D * 2;                 // 16 KB each for K and V
constexpr int AT_P_BYTES = AT_BM * AT_BN * 2;                 // 32 KB: two K-blocks of [128 rows][64 keys]
constexpr int AT_SMEM = AT_Q_BYTES + AT_STAGES * 2 * AT_KV_BYTES + AT_P_BYTES + 128 /*barriers*/;
constexpr int AT_THREADS = 192;
constexpr int AT_TMEM_COLS = 256;

struct AttnTcP {
  int B, H, Lq, Lk;
  int ldo, out_dtype;
  int use_band, start, end;
  float scale_log2;                 // softmax scale * log2(e)
  void* out;
  float* lse;
  const uint8_t* kmask;
  pka_dropout drop;
};

__device__ __forceinline__ float fast_exp2(float x) {       // MUFU.EX2 (2 ulp); exp2f() adds range fix-ups around it
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr uint32_t kIdescS = make_idesc(AT_BM, AT_BN);                  // S = Q K^T
constexpr uint32_t kIdescO = make_idesc(AT_BM, AT_D, false, true);      // O = P V  (V is MN-major)

// Chunks of 32 keys whose keys are all allowed (the common case away from the band edge and the padded tail) skip the
// per-element mask tests; the tile maximum is taken on the raw scores and scaled once (monotone for scale > 0).
// Two tensor-memory passes over the scores of a tile (maximum, then exponentials).  A single-pass instantiation that kept
// all 128 scores of a row in registers was measured and removed (round 2: 106 us vs 73 us at B=4, H=8, T=1600 -- 128 more
// live registers per thread cost more than the second tcgen05.ld pass saves).
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                   const __grid_constant__ CUtensorMap mapV, const AttnTcP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + AT_Q_BYTES;
  uint8_t* sV = sK + AT_STAGES * AT_KV_BYTES;
  uint8_t* sP = sV + AT_STAGES * AT_KV_BYTES;
  uint64_t* bars = (uint64_t*)(sP + AT_P_BYTES);
  // bars: [0] q_full, [1..2] kv_full, [3..4] kv_empty, [5] s_full, [6] p_full, [7] o_full; then the TMEM base slot
  uint32_t* tmem_slot = (uint32_t*)(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BM, h = blockIdx.y, b = blockIdx.z;

  if ((smem_u32(smem) & 1023u) != 0) __trap();     // the swizzle atoms need a 1 KB aligned base

  // key tiles this query tile can see
  int jlo = 0, jhi = p.Lk - 1;
  if (p.use_band) {
    jlo = max(0, q0 + p.start);
    jhi = min(p.Lk - 1, min(q0 + AT_BM - 1, p.Lq - 1) + p.end);
  }
  const int tile_lo = jlo / AT_BN;
  const int n_tiles = jhi >= jlo ? jhi / AT_BN - tile_lo + 1 : 0;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    for (int s = 0; s < AT_STAGES; ++s) { mbar_init(smem_u32(&bars[1 + s]), 1); mbar_init(smem_u32(&bars[3 + s]), 1); }
    mbar_init(smem_u32(&bars[5]), 1);
    mbar_init(smem_u32(&bars[6]), 128);
    mbar_init(smem_u32(&bars[7]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) tmem_alloc<AT_TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                    // the prologue above touched only shared / tensor memory
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

  if (warp == 4) {
    if (lane == 0 && n_tiles > 0) {                // ===== TMA producer
      tma_prefetch_desc(&mapQ); tma_prefetch_desc(&mapK); tma_prefetch_desc(&mapV);
      mbar_expect_tx(smem_u32(&bars[0]), AT_Q_BYTES);
      tma_load_3d(smem_u32(sQ), &mapQ, smem_u32(&bars[0]), h * AT_D, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % AT_STAGES, round = j / AT_STAGES;
        mbar_wait(smem_u32(&bars[3 + s]), (round & 1) ^ 1);
        const uint32_t full = smem_u32(&bars[1 + s]);
        mbar_expect_tx(full, 2 * AT_KV_BYTES);
        const int j0 = (tile_lo + j) * AT_BN;
        tma_load_3d(smem_u32(sK + s * AT_KV_BYTES), &mapK, full, h * AT_D, j0, b);
        tma_load_3d(smem_u32(sV + s * AT_KV_BYTES), &mapV, full, h * AT_D, j0, b);
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && n_tiles > 0) {                // ===== MMA issuer
      mbar_wait(smem_u32(&bars[0]), 0);
      const uint64_t dq = make_sdesc(smem_u32(sQ));
      const uint64_t dp = make_sdesc(smem_u32(sP));
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % AT_STAGES, round = j / AT_STAGES;
        mbar_wait(smem_u32(&bars[1 + s]), round & 1);
        tc_fence_after();
        const uint64_t dk = make_sdesc(smem_u32(sK + s * AT_KV_BYTES));
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)        // 32 B per 16-element K step inside the swizzle atom
          umma_f16(tmem_S, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), kIdescS, k ? 1u : 0u);
        umma_commit(smem_u32(&bars[5]));           // S_j ready
        mbar_wait(smem_u32(&bars[6]), j & 1);      // P_j in shared memory (and S_j, O_{j-1} consumed)
        tc_fence_after();
        const uint64_t dv = make_sdesc(smem_u32(sV + s * AT_KV_BYTES), AT_KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < AT_BN / 16; ++kk) {  // P: K-block (64 keys) = 16 KB, 32 B per step inside; V: 16 key rows = 2 KB
          const uint64_t da = dp + (uint64_t)((kk >> 2) * (AT_BM * 128 / 16) + (kk & 3) * 2);
          umma_f16(tmem_O, da, dv + (uint64_t)(kk * 128), kIdescO, (j | kk) ? 1u : 0u);   // O accumulates over all tiles
        }
        umma_commit(smem_u32(&bars[3 + s]));       // K_j / V_j stage free
      }
      umma_commit(smem_u32(&bars[7]));             // O complete
    }
  } else {                                         // ===== softmax warps 0..3: thread = query row
    const int r = warp * 32 + lane;
    const int i = q0 + r;
    const bool row_ok = i < p.Lq;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const uint8_t* km = p.kmask + (long long)b * p.Lk;
    const DropCtx dc = make_drop(p.drop);
    const unsigned long long drop_row = (((unsigned long long)b * p.H + h) * p.Lq + (row_ok ? i : 0)) * (unsigned long long)((p.Lk + 7) & ~7);
    float m_ref = -CUDART_INF_F, l_run = 0.f;      // exponent reference (>= every score seen minus 8), running sum
    uint8_t* prow = sP + (r >> 3) * 1024 + (r & 7) * 128;
    constexpr float kLazy = 8.f;                   // raise the reference only when a tile exceeds it by 2^8

    // allowed-key bit masks (key padding by warp ballot, band as a per-row bit range) and dropout keep bits of this
    // row for the 4 chunks of 32 keys of a tile.  Both depend on indices only, so the masks of tile j+1 are built while
    // the tensor core runs P_j V_j (the thread would otherwise just wait for the next scores).
    uint32_t allow[4], keep[4];
    auto tile_masks = [&](int jt, uint32_t (&al)[4], uint32_t (&kp)[4]) {
      const int jb = (tile_lo + jt) * AT_BN;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int jj = jb + c * 32 + lane;
        const uint32_t kbits = __ballot_sync(0xffffffffu, jj < p.Lk && km[jj] != 0);
        uint32_t bm = 0xffffffffu;
        if (p.use_band) {
          const int lo = max(0, i + p.start - (jb + c * 32)), hi = min(31, i + p.end - (jb + c * 32));
          bm = (hi >= lo) ? ((0xffffffffu >> (31 - hi)) & (0xffffffffu << lo)) : 0u;
        }
        al[c] = row_ok ? (kbits & bm) : 0u;
        kp[c] = 0xffffffffu;
        if (dc.p > 0.f && al[c] != 0u) {           // one Philox call per 8 consecutive keys (row pitch Lk8 % 8 == 0)
          uint32_t kb = 0u;
#pragma unroll
          for (int e = 0; e < 32; e += 8) kb |= dropout_bits8(dc, (drop_row + (unsigned long long)(jb + c * 32 + e)) >> 3) << e;
          kp[c] = kb;
        }
      }
    };
    if (n_tiles > 0) tile_masks(0, allow, keep);

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(smem_u32(&bars[5]), j & 1);        // S_j complete -- and with it every earlier MMA (P_{j-1} V_{j-1} included)
      tc_fence_after();
      // the maximum is taken chunk by chunk and the scores are loaded again for the exponentials (few live registers,
      // twice the tensor-memory reads)
      uint32_t sv[1][32];
      float t_max = -CUDART_INF_F;
      // tile maximum on the raw scores (scaled once: x -> x * c is monotone for c > 0)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld32(tmem_S + lane_addr + (uint32_t)(c * 32), sv[0]);
        const uint32_t (&x)[32] = sv[0];
        if (allow[c] == 0xffffffffu) {
#pragma unroll
          for (int e = 0; e < 32; ++e) t_max = fmaxf(t_max, __uint_as_float(x[e]));
        } else if (allow[c] != 0u) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if ((allow[c] >> e) & 1u) t_max = fmaxf(t_max, __uint_as_float(x[e]));
        }
      }
      t_max *= p.scale_log2;                       // scale_log2 > 0: -inf stays -inf
      // lazy reference update: first allowed key of the row, or a tile that exceeds the reference by more than 2^kLazy
      const bool first = (m_ref == -CUDART_INF_F) && (t_max != -CUDART_INF_F);
      const bool raise = (m_ref != -CUDART_INF_F) && (t_max > m_ref + kLazy);
      if (__any_sync(0xffffffffu, raise)) {        // rescale O (tensor memory) and l of the rows that moved; rare after tile 0
        const float alpha = raise ? fast_exp2(m_ref - t_max) : 1.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t ov[32];
          tmem_ld32(tmem_O + lane_addr + (uint32_t)(c * 32), ov);
#pragma unroll
          for (int e = 0; e < 32; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
          tmem_st32(tmem_O + lane_addr + (uint32_t)(c * 32), ov);
        }
        tmem_st_wait();
        l_run *= alpha;
      }
      if (first || raise) m_ref = t_max;
      const float m_use = (m_ref == -CUDART_INF_F) ? 0.f : m_ref;
      // probabilities -> bf16 P tile in shared memory (UMMA K-major, 128B swizzle)
      float l_tile = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float pv[32];
        tmem_ld32(tmem_S + lane_addr + (uint32_t)(c * 32), sv[0]);
        const uint32_t (&x)[32] = sv[0];
        if (allow[c] == 0xffffffffu) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float pe = fast_exp2(fmaf(__uint_as_float(x[e]), p.scale_log2, -m_use));
            l_tile += pe;
            pv[e] = pe;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float pe = ((allow[c] >> e) & 1u) ? fast_exp2(fmaf(__uint_as_float(x[e]), p.scale_log2, -m_use)) : 0.f;
            l_tile += pe;
            pv[e] = pe;
          }
        }
        if (dc.p > 0.f) {
#pragma unroll
          for (int e = 0; e < 32; ++e) pv[e] = ((keep[c] >> e) & 1u) ? pv[e] * dc.scale : 0.f;
        }
        uint8_t* pblk = prow + (c >> 1) * (AT_BM * 128);
#pragma unroll
        for (int g = 0; g < 4; ++g) {              // 8 keys = one 16-byte chunk
          uint4 pk;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(pv[g * 8 + 0], pv[g * 8 + 1]), h1 = __floats2bfloat162_rn(pv[g * 8 + 2], pv[g * 8 + 3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(pv[g * 8 + 4], pv[g * 8 + 5]), h3 = __floats2bfloat162_rn(pv[g * 8 + 6], pv[g * 8 + 7]);
          pk.x = *(uint32_t*)&h0; pk.y = *(uint32_t*)&h1; pk.z = *(uint32_t*)&h2; pk.w = *(uint32_t*)&h3;
          const int chunk = (c & 1) * 4 + g;
          *reinterpret_cast<uint4*>(pblk + ((chunk ^ (r & 7)) << 4)) = pk;
        }
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars[6]));             // P_j written, S_j consumed, O rescaled: P_j V_j (then S_{j+1}) may issue
      l_run += l_tile;
      if (j + 1 < n_tiles) tile_masks(j + 1, allow, keep);       // overlaps the P_j V_j and Q K_{j+1} MMAs
    }
    // tcgen05.ld is warp-collective (.sync.aligned): every lane takes part, rows beyond Lq just do not store
    float o[AT_D];
    if (n_tiles > 0) {
      mbar_wait(smem_u32(&bars[7]), 0);            // every P_j V_j has been accumulated
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t ov[32];
        tmem_ld32(tmem_O + lane_addr + (uint32_t)(c * 32), ov);
#pragma unroll
        for (int e = 0; e < 32; ++e) o[c * 32 + e] = __uint_as_float(ov[e]);
      }
    } else {
#pragma unroll
      for (int d = 0; d < AT_D; ++d) o[d] = 0.f;
    }
    if (row_ok) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      if (p.out_dtype == PKA_BF16) {
        __nv_bfloat16* dst = (__nv_bfloat16*)p.out + ((long long)b * p.Lq + i) * p.ldo + h * AT_D;
#pragma unroll
        for (int d = 0; d < AT_D; d += 8) {
          uint4 pk;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(o[d] * inv, o[d + 1] * inv), h1 = __floats2bfloat162_rn(o[d + 2] * inv, o[d + 3] * inv);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(o[d + 4] * inv, o[d + 5] * inv), h3 = __floats2bfloat162_rn(o[d + 6] * inv, o[d + 7] * inv);
          pk.x = *(uint32_t*)&h0; pk.y = *(uint32_t*)&h1; pk.z = *(uint32_t*)&h2; pk.w = *(uint32_t*)&h3;
          *reinterpret_cast<uint4*>(dst + d) = pk;
        }
      } else {
        float* dst = (float*)p.out + ((long long)b * p.Lq + i) * p.ldo + h * AT_D;
#pragma unroll
        for (int d = 0; d < AT_D; d += 4)
          *reinterpret_cast<float4*>(dst + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
      }
      // natural-log lse of the scaled scores, as the SIMT kernels store it
      p.lse[((long long)b * p.H + h) * p.Lq + i] = l_run > 0.f ? (m_ref + log2f(l_run)) * 0.69314718055994531f : -CUDART_INF_F;
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<AT_TMEM_COLS>(tmem_base);
}

}  // namespace pka

using namespace pka;

extern "C" int pka_attn_tc_fwd(const pka_attn_desc* d, const void* q, const void* k, const void* v, const uint8_t* key_mask,
                               void* out, int out_dtype, float* lse, void* stream) {
  PKA_REQUIRE(d && q && k && v && key_mask && out && lse, PKA_EINVAL, "attn_tc_fwd: null pointer");
  PKA_REQUIRE(d->B > 0 && d->H > 0 && d->Lq > 0 && d->Lk > 0, PKA_EINVAL, "attn_tc_fwd: bad sizes B=%d H=%d Lq=%d Lk=%d", d->B, d->H, d->Lq, d->Lk);
  PKA_REQUIRE(d->dk == AT_D && d->dv == AT_D, PKA_EUNSUPPORTED, "attn_tc_fwd: head dim %d/%d (tensor-core path is built for 64)", d->dk, d->dv);
  PKA_REQUIRE(d->B <= 65535 && d->H <= 65535, PKA_EUNSUPPORTED, "attn_tc_fwd: B or H exceeds grid limits");
  PKA_REQUIRE(out_dtype == PKA_BF16 || out_dtype == PKA_F32, PKA_EUNSUPPORTED, "attn_tc_fwd: out dtype %d", out_dtype);
  PKA_REQUIRE((out_dtype == PKA_BF16 ? d->ldo % 8 : d->ldo % 4) == 0 && aligned16(out), PKA_EALIGN, "attn_tc_fwd: out must be 16-byte aligned rows");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    PKA_REQUIRE(e == cudaSuccess, PKA_ELAUNCH, "attn_tc_fwd: cannot opt in to %d bytes of shared memory: %s", AT_SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  CUtensorMap mapQ, mapK, mapV;
  const uint64_t cols = (uint64_t)d->H * AT_D;
  int rc = make_map(&mapQ, q, cols, d->Lq, d->B, (uint64_t)d->ldq * 2, (uint64_t)d->Lq * d->ldq * 2, AT_BM, 1, "attn_tc Q");
  if (rc) return rc;
  rc = make_map(&mapK, k, cols, d->Lk, d->B, (uint64_t)d->ldk * 2, (uint64_t)d->Lk * d->ldk * 2, AT_BN, 1, "attn_tc K");
  if (rc) return rc;
  rc = make_map(&mapV, v, cols, d->Lk, d->B, (uint64_t)d->ldv * 2, (uint64_t)d->Lk * d->ldv * 2, AT_BN, 1, "attn_tc V");
  if (rc) return rc;
  AttnTcP p;
  p.B = d->B; p.H = d->H; p.Lq = d->Lq; p.Lk = d->Lk;
  p.ldo = d->ldo; p.out_dtype = out_dtype;
  p.use_band = d->use_band; p.start = d->band_start; p.end = d->band_end;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.out = out; p.lse = lse; p.kmask = key_mask; p.drop = d->drop;
  dim3 grid((d->Lq + AT_BM - 1) / AT_BM, d->H, d->B);
  launch_k(attn_tc_fwd_kernel, grid, AT_THREADS, AT_SMEM, as_stream(stream), mapQ, mapK, mapV, p);
  return check_launch("attn_tc_fwd");
}

namespace pka {

// =====================================================================================================================
// Backward on the tensor cores.  One kernel template, two launches (deterministic: no atomics):
//   MODE 0 (dQ):    rows = 128 queries of one (utterance, head); 64-key tiles stream by.      dQ  = scale * dS  K
//   MODE 1 (dK,dV): rows = 128 keys;                             64-query tiles stream by.    dK  = scale * dS^T Q,  dV = (P.M)^T dO
// Per streamed tile (R1,R2 = resident row operands, C1,C2 = streamed ones):
//   S  = R1 C1^T, dP = R2 C2^T        tcgen05.mma 128x64x16 x4 each (all operands K-major)      -> TMEM cols [0,64), [64,128)
//   threads (one per row) recompute P = exp2(S*c - lse), apply the mask predicate and the Philox keep bits, form
//   dS = P (M dP - delta) scale and write dS (and P.M in MODE 1) as bf16 UMMA K-major tiles into shared memory
//   acc0 += dS C1, acc1 += (P.M) C2   tcgen05.mma 128x64x16 x4 each, B = the same streamed tile read MN-major,
//                                     accumulating in TMEM cols [128,192), [192,256) across all tiles
// MODE 0 also produces delta_i = dO_i . O_i (read by MODE 1, which is launched after it on the same stream).
constexpr int AB_BM = 128, AB_BN = 64, AB_STAGES = 2;
constexpr int AB_R_BYTES = AB_BM * AT_D * 2;                  // 16 KB
constexpr int AB_C_BYTES = AB_BN * AT_D * 2;                  // 8 KB
constexpr int AB_E_BYTES = AB_BM * AB_BN * 2;                 // 16 KB
constexpr int AB_SMEM = 2 * AB_R_BYTES + AB_STAGES * 2 * AB_C_BYTES + 2 * AB_E_BYTES + 128 /*barriers*/ + 4 * AB_BN * 4 /*lse, delta*/;
constexpr int AB_TMEM_COLS = 256;

struct AttnBwdP {
  int B, H, Lq, Lk;
  int use_band, start, end;
  float scale, scale_log2;
  const uint8_t* kmask;
  const float* lse;
  float* delta;
  const __nv_bfloat16* o; const __nv_bfloat16* dout; int ldo;      // MODE 0: delta = dO . O
  __nv_bfloat16* g0; int ldg0;                                      // MODE 0: dq; MODE 1: dk
  __nv_bfloat16* g1; int ldg1;                                      // MODE 1: dv
  pka_dropout drop;
};

constexpr uint32_t kIdescSB = make_idesc(AB_BM, AB_BN);                 // S, dP: both operands K-major
constexpr uint32_t kIdescAB = make_idesc(AB_BM, AT_D, false, true);     // accumulators: B MN-major

__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, const float (&v)[AT_D], float mul) {
#pragma unroll
  for (int d = 0; d < AT_D; d += 8) {
    uint4 pk;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[d] * mul, v[d + 1] * mul), h1 = __floats2bfloat162_rn(v[d + 2] * mul, v[d + 3] * mul);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[d + 4] * mul, v[d + 5] * mul), h3 = __floats2bfloat162_rn(v[d + 6] * mul, v[d + 7] * mul);
    pk.x = *(uint32_t*)&h0; pk.y = *(uint32_t*)&h1; pk.z = *(uint32_t*)&h2; pk.w = *(uint32_t*)&h3;
    *reinterpret_cast<uint4*>(dst + d) = pk;
  }
}

template <int MODE>
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap mapR1, const __grid_constant__ CUtensorMap mapR2,
                   const __grid_constant__ CUtensorMap mapC1, const __grid_constant__ CUtensorMap mapC2, const AttnBwdP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sR1 = smem;
  uint8_t* sR2 = smem + AB_R_BYTES;
  uint8_t* sC1 = sR2 + AB_R_BYTES;
  uint8_t* sC2 = sC1 + AB_STAGES * AB_C_BYTES;
  uint8_t* sE1 = sC2 + AB_STAGES * AB_C_BYTES;
  uint8_t* sE2 = sE1 + AB_E_BYTES;
  uint64_t* bars = (uint64_t*)(sE2 + AB_E_BYTES);
  // bars: [0] r_full, [1..2] c_full, [3..4] c_empty, [5] sp_full, [6] e_full (128 arrivals), [7] acc_done (E free / final)
  uint32_t* tmem_slot = (uint32_t*)(bars + 8);
  float* col_lse = (float*)(bars + 16);            // [2][64] (MODE 1)
  float* col_del = col_lse + 2 * AB_BN;            // [2][64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * AB_BM, h = blockIdx.y, b = blockIdx.z;
  const int Lrow = MODE == 0 ? p.Lq : p.Lk, Lcol = MODE == 0 ? p.Lk : p.Lq;

  if ((smem_u32(smem) & 1023u) != 0) __trap();

  int clo = 0, chi = Lcol - 1;
  if (p.use_band) {
    const int xl = min(x0 + AB_BM - 1, Lrow - 1);
    if (MODE == 0) { clo = max(0, x0 + p.start); chi = min(Lcol - 1, xl + p.end); }        // keys visible to these queries
    else { clo = max(0, x0 - p.end); chi = min(Lcol - 1, xl - p.start); }                   // queries that see these keys
  }
  const int tile_lo = clo / AB_BN;
  const int n_tiles = chi >= clo ? chi / AB_BN - tile_lo + 1 : 0;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    for (int s = 0; s < AB_STAGES; ++s) { mbar_init(smem_u32(&bars[1 + s]), 1); mbar_init(smem_u32(&bars[3 + s]), 1); }
    mbar_init(smem_u32(&bars[5]), 1);
    mbar_init(smem_u32(&bars[6]), 128);
    mbar_init(smem_u32(&bars[7]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) tmem_alloc<AB_TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                    // the prologue above touched only shared / tensor memory
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 64, tmem_A0 = tmem_base + 128, tmem_A1 = tmem_base + 192;

  if (warp == 4) {
    if (lane == 0 && n_tiles > 0) {                // ===== TMA producer
      tma_prefetch_desc(&mapR1); tma_prefetch_desc(&mapR2); tma_prefetch_desc(&mapC1); tma_prefetch_desc(&mapC2);
      mbar_expect_tx(smem_u32(&bars[0]), 2 * AB_R_BYTES);
      tma_load_3d(smem_u32(sR1), &mapR1, smem_u32(&bars[0]), h * AT_D, x0, b);
      tma_load_3d(smem_u32(sR2), &mapR2, smem_u32(&bars[0]), h * AT_D, x0, b);
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % AB_STAGES, round = t / AB_STAGES;
        mbar_wait(smem_u32(&bars[3 + s]), (round & 1) ^ 1);
        const uint32_t full = smem_u32(&bars[1 + s]);
        mbar_expect_tx(full, 2 * AB_C_BYTES);
        const int c0 = (tile_lo + t) * AB_BN;
        tma_load_3d(smem_u32(sC1 + s * AB_C_BYTES), &mapC1, full, h * AT_D, c0, b);
        tma_load_3d(smem_u32(sC2 + s * AB_C_BYTES), &mapC2, full, h * AT_D, c0, b);
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && n_tiles > 0) {                // ===== MMA issuer
      mbar_wait(smem_u32(&bars[0]), 0);
      const uint64_t dr1 = make_sdesc(smem_u32(sR1)), dr2 = make_sdesc(smem_u32(sR2));
      const uint64_t de1 = make_sdesc(smem_u32(sE1)), de2 = make_sdesc(smem_u32(sE2));
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % AB_STAGES, round = t / AB_STAGES;
        mbar_wait(smem_u32(&bars[1 + s]), round & 1);
        tc_fence_after();
        const uint32_t c1 = smem_u32(sC1 + s * AB_C_BYTES), c2 = smem_u32(sC2 + s * AB_C_BYTES);
        const uint64_t dc1 = make_sdesc(c1), dc2 = make_sdesc(c2);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) umma_f16(tmem_S, dr1 + (uint64_t)(k * 2), dc1 + (uint64_t)(k * 2), kIdescSB, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) umma_f16(tmem_dP, dr2 + (uint64_t)(k * 2), dc2 + (uint64_t)(k * 2), kIdescSB, k ? 1u : 0u);
        umma_commit(smem_u32(&bars[5]));           // S_t, dP_t ready
        mbar_wait(smem_u32(&bars[6]), t & 1);      // dS_t (and P.M_t) in shared memory
        tc_fence_after();
        const uint64_t dm1 = make_sdesc(c1, AB_C_BYTES), dm2 = make_sdesc(c2, AB_C_BYTES);
#pragma unroll
        for (int k = 0; k < AB_BN / 16; ++k)       // reduction over the 64 streamed rows: 16 rows = 2 KB of the MN-major tile
          umma_f16(tmem_A0, de1 + (uint64_t)(k * 2), dm1 + (uint64_t)(k * 128), kIdescAB, (t | k) ? 1u : 0u);
        if (MODE == 1) {
#pragma unroll
          for (int k = 0; k < AB_BN / 16; ++k)
            umma_f16(tmem_A1, de2 + (uint64_t)(k * 2), dm2 + (uint64_t)(k * 128), kIdescAB, (t | k) ? 1u : 0u);
        }
        umma_commit(smem_u32(&bars[7]));           // E buffers free again / accumulators complete after the last tile
        umma_commit(smem_u32(&bars[3 + s]));       // streamed stage free
      }
    }
  } else {                                         // ===== compute warps 0..3: thread = row
    const int r = warp * 32 + lane;
    const int x = x0 + r;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const uint8_t* km = p.kmask + (long long)b * p.Lk;
    const DropCtx dc = make_drop(p.drop);
    const unsigned long long Lk8 = (unsigned long long)((p.Lk + 7) & ~7);
    const unsigned long long bh = (unsigned long long)b * p.H + h;
    bool row_ok = x < Lrow;
    float row_lse2 = 0.f, row_del = 0.f;
    if (MODE == 0) {
      if (row_ok) {
        const float l = p.lse[bh * p.Lq + x];
        const uint4* po = reinterpret_cast<const uint4*>(p.o + ((long long)b * p.Lq + x) * p.ldo + h * AT_D);
        const uint4* pd = reinterpret_cast<const uint4*>(p.dout + ((long long)b * p.Lq + x) * p.ldo + h * AT_D);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < AT_D / 8; ++c) {
          const uint4 a = po[c], g = pd[c];
          const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
          const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 fa = __bfloat1622float2(ah[e]), fg = __bfloat1622float2(gh[e]);
            acc = fmaf(fa.x, fg.x, acc); acc = fmaf(fa.y, fg.y, acc);
          }
        }
        row_del = acc;
        p.delta[bh * p.Lq + x] = acc;
        row_lse2 = l * 1.4426950408889634f;
        if (l == -CUDART_INF_F) row_ok = false;
      }
    } else {
      row_ok = row_ok && km[x] != 0;
    }
    uint8_t* e1row = sE1 + (r >> 3) * 1024 + (r & 7) * 128;
    uint8_t* e2row = sE2 + (r >> 3) * 1024 + (r & 7) * 128;

    for (int t = 0; t < n_tiles; ++t) {
      const int c0 = (tile_lo + t) * AB_BN;
      uint32_t allow[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int cb = c0 + c * 32;
        uint32_t bits;
        int lo = 0, hi = 31;
        if (MODE == 0) {
          const int jj = cb + lane;
          bits = __ballot_sync(0xffffffffu, jj < p.Lk && km[jj] != 0);
          if (p.use_band) { lo = max(0, x + p.start - cb); hi = min(31, x + p.end - cb); }
        } else {
          bits = 0xffffffffu;
          hi = min(31, p.Lq - 1 - cb);
          if (p.use_band) { lo = max(0, x - p.end - cb); hi = min(hi, x - p.start - cb); }
        }
        const uint32_t bm = (hi >= lo) ? ((0xffffffffu >> (31 - hi)) & (0xffffffffu << lo)) : 0u;
        allow[c] = row_ok ? (bits & bm) : 0u;
      }
      if (MODE == 1) {                             // per-query statistics of this tile -> shared memory (parity buffer)
        if (r < AB_BN) {
          const int i = c0 + r;
          float l = CUDART_INF_F, dl = 0.f;        // +inf: exp2(s - inf) = 0 for dead / out-of-range queries
          if (i < p.Lq) {
            const float lv = p.lse[bh * p.Lq + i];
            if (lv != -CUDART_INF_F) l = lv * 1.4426950408889634f;
            dl = p.delta[bh * p.Lq + i];
          }
          col_lse[(t & 1) * AB_BN + r] = l;
          col_del[(t & 1) * AB_BN + r] = dl;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      mbar_wait(smem_u32(&bars[5]), t & 1);
      tc_fence_after();
      if (t > 0) mbar_wait(smem_u32(&bars[7]), (t - 1) & 1);      // the MMAs of tile t-1 have finished reading E
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld32_nowait(tmem_S + lane_addr + (uint32_t)(c * 32), sv);
        tmem_ld32_nowait(tmem_dP + lane_addr + (uint32_t)(c * 32), dv);
        tmem_ld_wait();
        float ds[32], pm[32];
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          // keep bits of this thread's 8 (row, column) pairs
          uint32_t kb = 0xffu;
          if (dc.p > 0.f) {
            if (MODE == 0) {                       // row = query: 8 consecutive keys share one Philox call
              kb = dropout_bits8(dc, ((bh * p.Lq + x) * Lk8 + (unsigned long long)(c0 + c * 32 + e)) >> 3);
            } else {
              // row = key x, columns = 8 consecutive queries.  A call covers 8 consecutive KEYS of one query, i.e. the 8
              // lanes of an aligned lane group: lane L evaluates the call of query (e + (L & 7)) for its group's keys
              // and the group exchanges the result bytes (8 shuffles instead of 8 Philox calls per thread).
              const int qi = c0 + c * 32 + e + (lane & 7);
              const uint32_t mine = dropout_bits8(dc, ((bh * p.Lq + (unsigned long long)qi) * Lk8 + (unsigned long long)(x & ~7)) >> 3);
              kb = 0u;
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const uint32_t other = __shfl_sync(0xffffffffu, mine, (lane & ~7) | u);     // byte of query e+u, keys of my group
                kb |= ((other >> (lane & 7)) & 1u) << u;
              }
            }
          }
          // per-column statistics of the 8 columns (MODE 1: two 16-byte broadcast reads instead of 16 scalar ones)
          float l2v[8], dlv[8];
          if (MODE == 0) {
#pragma unroll
            for (int u = 0; u < 8; ++u) { l2v[u] = row_lse2; dlv[u] = row_del; }
          } else {
            const float4* pl = reinterpret_cast<const float4*>(col_lse + (t & 1) * AB_BN + c * 32 + e);
            const float4* pd = reinterpret_cast<const float4*>(col_del + (t & 1) * AB_BN + c * 32 + e);
            const float4 la = pl[0], lb = pl[1], da = pd[0], db = pd[1];
            l2v[0] = la.x; l2v[1] = la.y; l2v[2] = la.z; l2v[3] = la.w; l2v[4] = lb.x; l2v[5] = lb.y; l2v[6] = lb.z; l2v[7] = lb.w;
            dlv[0] = da.x; dlv[1] = da.y; dlv[2] = da.z; dlv[3] = da.w; dlv[4] = db.x; dlv[5] = db.y; dlv[6] = db.z; dlv[7] = db.w;
          }
          const uint32_t al8 = (allow[c] >> e) & 0xffu;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            // MUFU.EX2 directly: exp2(-inf) = +0 covers dead / out-of-range columns (lse = +inf there)
            float pe = fast_exp2(fmaf(__uint_as_float(sv[e + u]), p.scale_log2, -l2v[u]));
            if (al8 != 0xffu) pe = ((al8 >> u) & 1u) ? pe : 0.f;        // groups of 8 fully allowed keys skip the select
            const float pmu = (dc.p > 0.f) ? (((kb >> u) & 1u) ? pe * dc.scale : 0.f) : pe;
            pm[e + u] = pmu;
            ds[e + u] = (pmu * __uint_as_float(dv[e + u]) - pe * dlv[u]) * p.scale;
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk = ((c * 4 + g) ^ (r & 7)) << 4;
          uint4 pk;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(ds[g * 8 + 0], ds[g * 8 + 1]), h1 = __floats2bfloat162_rn(ds[g * 8 + 2], ds[g * 8 + 3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(ds[g * 8 + 4], ds[g * 8 + 5]), h3 = __floats2bfloat162_rn(ds[g * 8 + 6], ds[g * 8 + 7]);
          pk.x = *(uint32_t*)&h0; pk.y = *(uint32_t*)&h1; pk.z = *(uint32_t*)&h2; pk.w = *(uint32_t*)&h3;
          *reinterpret_cast<uint4*>(e1row + chunk) = pk;
          if (MODE == 1) {
            __nv_bfloat162 g0 = __floats2bfloat162_rn(pm[g * 8 + 0], pm[g * 8 + 1]), g1 = __floats2bfloat162_rn(pm[g * 8 + 2], pm[g * 8 + 3]);
            __nv_bfloat162 g2 = __floats2bfloat162_rn(pm[g * 8 + 4], pm[g * 8 + 5]), g3 = __floats2bfloat162_rn(pm[g * 8 + 6], pm[g * 8 + 7]);
            pk.x = *(uint32_t*)&g0; pk.y = *(uint32_t*)&g1; pk.z = *(uint32_t*)&g2; pk.w = *(uint32_t*)&g3;
            *reinterpret_cast<uint4*>(e2row + chunk) = pk;
          }
        }
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars[6]));
    }
    // ---- epilogue: accumulators -> bf16 rows
    float v[AT_D];
    if (n_tiles > 0) {
      mbar_wait(smem_u32(&bars[7]), (n_tiles - 1) & 1);
      tc_fence_after();
    }
    const bool wr = x < Lrow;
    const long long orow = (long long)b * Lrow + x;
#pragma unroll
    for (int a = 0; a < (MODE == 0 ? 1 : 2); ++a) {
      if (n_tiles > 0) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t av[32];
          tmem_ld32((a == 0 ? tmem_A0 : tmem_A1) + lane_addr + (uint32_t)(c * 32), av);
#pragma unroll
          for (int e = 0; e < 32; ++e) v[c * 32 + e] = __uint_as_float(av[e]);
        }
      } else {
#pragma unroll
        for (int d = 0; d < AT_D; ++d) v[d] = 0.f;
      }
      if (wr) {
        __nv_bfloat16* dst = a == 0 ? p.g0 + orow * p.ldg0 + h * AT_D : p.g1 + orow * p.ldg1 + h * AT_D;
        store_row_bf16(dst, v, 1.f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<AB_TMEM_COLS>(tmem_base);
}

}  // namespace pka

extern "C" int pka_attn_tc_bwd(const pka_attn_desc* d, const void* q, const void* k, const void* v, const uint8_t* key_mask,
                               const void* out, const void* dout, const float* lse, float* delta_ws, void* dq, void* dk,
                               void* dv, void* stream) {
  using namespace pka;
  PKA_REQUIRE(d && q && k && v && key_mask && out && dout && lse && delta_ws && dq && dk && dv, PKA_EINVAL, "attn_tc_bwd: null pointer");
  PKA_REQUIRE(d->B > 0 && d->H > 0 && d->Lq > 0 && d->Lk > 0, PKA_EINVAL, "attn_tc_bwd: bad sizes");
  PKA_REQUIRE(d->dk == AT_D && d->dv == AT_D, PKA_EUNSUPPORTED, "attn_tc_bwd: head dim %d/%d (tensor-core path is built for 64)", d->dk, d->dv);
  PKA_REQUIRE(d->B <= 65535 && d->H <= 65535, PKA_EUNSUPPORTED, "attn_tc_bwd: B or H exceeds grid limits");
  PKA_REQUIRE(d->ldo % 8 == 0 && d->ldq % 8 == 0 && d->ldk % 8 == 0 && d->ldv % 8 == 0 && aligned16(out) && aligned16(dout) &&
              aligned16(dq) && aligned16(dk) && aligned16(dv), PKA_EALIGN, "attn_tc_bwd: rows must be 16-byte aligned");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
    PKA_REQUIRE(e == cudaSuccess, PKA_ELAUNCH, "attn_tc_bwd: cannot opt in to %d bytes of shared memory: %s", AB_SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  const uint64_t cols = (uint64_t)d->H * AT_D;
  CUtensorMap mQr, mDOr, mKc, mVc, mKr, mVr, mQc, mDOc;
  int rc;
  if ((rc = make_map(&mQr, q, cols, d->Lq, d->B, (uint64_t)d->ldq * 2, (uint64_t)d->Lq * d->ldq * 2, AB_BM, 1, "attn_tc_bwd Q"))) return rc;
  if ((rc = make_map(&mDOr, dout, cols, d->Lq, d->B, (uint64_t)d->ldo * 2, (uint64_t)d->Lq * d->ldo * 2, AB_BM, 1, "attn_tc_bwd dO"))) return rc;
  if ((rc = make_map(&mKc, k, cols, d->Lk, d->B, (uint64_t)d->ldk * 2, (uint64_t)d->Lk * d->ldk * 2, AB_BN, 1, "attn_tc_bwd K"))) return rc;
  if ((rc = make_map(&mVc, v, cols, d->Lk, d->B, (uint64_t)d->ldv * 2, (uint64_t)d->Lk * d->ldv * 2, AB_BN, 1, "attn_tc_bwd V"))) return rc;
  if ((rc = make_map(&mKr, k, cols, d->Lk, d->B, (uint64_t)d->ldk * 2, (uint64_t)d->Lk * d->ldk * 2, AB_BM, 1, "attn_tc_bwd K"))) return rc;
  if ((rc = make_map(&mVr, v, cols, d->Lk, d->B, (uint64_t)d->ldv * 2, (uint64_t)d->Lk * d->ldv * 2, AB_BM, 1, "attn_tc_bwd V"))) return rc;
  if ((rc = make_map(&mQc, q, cols, d->Lq, d->B, (uint64_t)d->ldq * 2, (uint64_t)d->Lq * d->ldq * 2, AB_BN, 1, "attn_tc_bwd Q"))) return rc;
  if ((rc = make_map(&mDOc, dout, cols, d->Lq, d->B, (uint64_t)d->ldo * 2, (uint64_t)d->Lq * d->ldo * 2, AB_BN, 1, "attn_tc_bwd dO"))) return rc;
  AttnBwdP p;
  p.B = d->B; p.H = d->H; p.Lq = d->Lq; p.Lk = d->Lk;
  p.use_band = d->use_band; p.start = d->band_start; p.end = d->band_end;
  p.scale = d->scale; p.scale_log2 = d->scale * 1.4426950408889634f;
  p.kmask = key_mask; p.lse = lse; p.delta = delta_ws;
  p.o = (const __nv_bfloat16*)out; p.dout = (const __nv_bfloat16*)dout; p.ldo = d->ldo;
  p.drop = d->drop;
  p.g0 = (__nv_bfloat16*)dq; p.ldg0 = d->ldq; p.g1 = nullptr; p.ldg1 = 0;
  dim3 gq((d->Lq + AB_BM - 1) / AB_BM, d->H, d->B);
  launch_k(attn_tc_bwd_kernel<0>, gq, AT_THREADS, AB_SMEM, as_stream(stream), mQr, mDOr, mKc, mVc, p);
  rc = check_launch("attn_tc_bwd(dq)");
  if (rc) return rc;
  p.g0 = (__nv_bfloat16*)dk; p.ldg0 = d->ldk; p.g1 = (__nv_bfloat16*)dv; p.ldg1 = d->ldv;
  dim3 gk((d->Lk + AB_BM - 1) / AB_BM, d->H, d->B);
  launch_k(attn_tc_bwd_kernel<1>, gk, AT_THREADS, AB_SMEM, as_stream(stream), mKr, mVr, mQc, mDOc, p);
  return check_launch("attn_tc_bwd(dkv)");
}
