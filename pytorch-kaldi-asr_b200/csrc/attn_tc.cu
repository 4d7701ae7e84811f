// Banded / key-padded multi-head attention on the 5th-generation tensor cores (sm_100a), forward pass.
// replaces: ScaledDotProductAttention.forward (T/Modules.py:75-97) + the masks of T/Models.py:27-49 + the head
// replication of T/SubLayers.py:49-59, for bf16 activations with head dim 64 (the only head dim the reference's
// Transformer can build: T/Models.py:250-251 never forwards d_k/d_v to the Decoder).
//
// One CTA = one (utterance, head, 128-query tile); key tiles of 128 stream through a 2-stage TMA ring:
//   warp 4 lane 0   TMA producer: Q once, then K_j / V_j boxes [128 keys][64] (128B swizzle, OOB rows zero-filled)
//   warp 5 lane 0   MMA issuer:   S = Q K_j^T   (tcgen05.mma 128x128x16 x4, both operands K-major)  -> TMEM cols [0,128)
//                                 O_j = P_j V_j (tcgen05.mma 128x64x16  x8, P K-major from smem, V MN-major) -> TMEM cols [128,192)
//   warps 0-3       online softmax, one thread per query row (TMEM lane = row): tcgen05.ld S, mask predicate
//                   (key padding via warp ballot of the u8 mask, band via per-row bit range), running max / sum in the
//                   exp2 domain, P_j written to shared memory as bf16 in the UMMA K-major 128B-swizzle layout, then
//                   O += alpha-rescaled accumulation in registers from the TMEM partial product.
// Two CTAs are resident per SM (112 KB smem, 256 TMEM columns each), so one CTA's softmax overlaps the other's MMAs.
// The [B*H, Lq, Lk] probability tensor never exists; rows without an allowed key give 0 output and lse = -inf.
#include "tc_common.cuh"
#include <math_constants.h>

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

constexpr uint32_t kIdescS = make_idesc(AT_BM, AT_BN);                  // S = Q K^T
constexpr uint32_t kIdescO = make_idesc(AT_BM, AT_D, false, true);      // O = P V  (V is MN-major)

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
          umma_f16(tmem_O, da, dv + (uint64_t)(kk * 128), kIdescO, kk ? 1u : 0u);
        }
        umma_commit(smem_u32(&bars[7]));           // O_j partial product ready
        umma_commit(smem_u32(&bars[3 + s]));       // K_j / V_j stage free
      }
    }
  } else {                                         // ===== softmax warps 0..3: thread = query row
    const int r = warp * 32 + lane;
    const int i = q0 + r;
    const bool row_ok = i < p.Lq;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const uint8_t* km = p.kmask + (long long)b * p.Lk;
    const DropCtx dc = make_drop(p.drop);
    const unsigned long long drop_row = (((unsigned long long)b * p.H + h) * p.Lq + (row_ok ? i : 0)) * (unsigned long long)p.Lk;
    float m_run = -CUDART_INF_F, l_run = 0.f;
    float o[AT_D];
#pragma unroll
    for (int d = 0; d < AT_D; ++d) o[d] = 0.f;
    uint8_t* prow = sP + (r >> 3) * 1024 + (r & 7) * 128;

    for (int j = 0; j < n_tiles; ++j) {
      const int j0 = (tile_lo + j) * AT_BN;
      // allowed-key bit masks of this row for the 4 chunks of 32 keys
      uint32_t allow[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int jj = j0 + c * 32 + lane;
        const uint32_t kbits = __ballot_sync(0xffffffffu, jj < p.Lk && km[jj] != 0);
        uint32_t bm = 0xffffffffu;
        if (p.use_band) {
          const int lo = max(0, i + p.start - (j0 + c * 32)), hi = min(31, i + p.end - (j0 + c * 32));
          bm = (hi >= lo) ? ((0xffffffffu >> (31 - hi)) & (0xffffffffu << lo)) : 0u;
        }
        allow[c] = row_ok ? (kbits & bm) : 0u;
      }
      mbar_wait(smem_u32(&bars[5]), j & 1);
      tc_fence_after();
      // pass 1: tile maximum of the scaled scores (exp2 domain)
      float t_max = -CUDART_INF_F;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32];
        tmem_ld32(tmem_S + lane_addr + (uint32_t)(c * 32), sv);
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if ((allow[c] >> e) & 1u) t_max = fmaxf(t_max, __uint_as_float(sv[e]) * p.scale_log2);
      }
      const float m_new = fmaxf(m_run, t_max);
      const float m_use = (m_new == -CUDART_INF_F) ? 0.f : m_new;
      const float alpha = (m_run == -CUDART_INF_F) ? 0.f : exp2f(m_run - m_new);
      // pass 2: probabilities -> bf16 P tile in shared memory (UMMA K-major, 128B swizzle)
      float l_tile = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32];
        tmem_ld32(tmem_S + lane_addr + (uint32_t)(c * 32), sv);
        float pv[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float pe = ((allow[c] >> e) & 1u) ? exp2f(fmaf(__uint_as_float(sv[e]), p.scale_log2, -m_use)) : 0.f;
          l_tile += pe;
          pv[e] = pe;
        }
        if (dc.p > 0.f) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if ((allow[c] >> e) & 1u)
              pv[e] = dropout_keep(dc, drop_row + (unsigned long long)(j0 + c * 32 + e)) ? pv[e] * dc.scale : 0.f;
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
      mbar_arrive(smem_u32(&bars[6]));
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      // O = alpha * O + P_j V_j
      mbar_wait(smem_u32(&bars[7]), j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t ov[32];
        tmem_ld32(tmem_O + lane_addr + (uint32_t)(c * 32), ov);
#pragma unroll
        for (int e = 0; e < 32; ++e) o[c * 32 + e] = fmaf(o[c * 32 + e], alpha, __uint_as_float(ov[e]));
      }
    }
    tc_fence_before();
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
      p.lse[((long long)b * p.H + h) * p.Lq + i] = l_run > 0.f ? (m_run + log2f(l_run)) * 0.69314718055994531f : -CUDART_INF_F;
    }
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
  attn_tc_fwd_kernel<<<grid, AT_THREADS, AT_SMEM, as_stream(stream)>>>(mapQ, mapK, mapV, p);
  return check_launch("attn_tc_fwd");
}
