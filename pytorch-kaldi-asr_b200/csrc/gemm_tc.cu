// bf16 tensor-core GEMM family for sm_100a: tcgen05.mma (UMMA 128x128x16, cta_group::1) with fp32 accumulators in
// TMEM, operands staged by TMA (cp.async.bulk.tensor, 128B swizzle) through a 3-stage mbarrier ring (two CTAs per SM), warp-specialised:
//   warp 0    TMA producer (one elected lane)
//   warp 1    MMA issuer   (one elected lane; tcgen05.commit releases smem stages / publishes the accumulator)
//   warp 2    TMEM allocator
//   warps 2-5 epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> bias / ReLU / dropout -> global stores
//
// Frame splicing (reference ConcatLayer, L/pytorch/TDNN.py:20-28) costs nothing here: the activation tensor is
// described to TMA as a 3-D tensor [utterance, frame, feature]; context c of a TDNN layer is the same box shifted by
// ctx[c] frames, and TMA's out-of-bounds ZERO FILL reproduces ConcatLayer's zero padding at the tensor edges.  The
// n_ctx shifted boxes are n_ctx K-segments accumulating into one TMEM tile, so [B,T,n_ctx*D] never exists.
//
// mode 0 ("rows"):   C[b,t,:] = epi( sum_seg A[b, t+shift[seg], :] . W[:, seg]^T )       forward and data-gradient
//                    one CTA per (utterance, 128-frame tile, 128-column tile); optional transposed copy Ct[n, b, t]
//                    (coalesced for free: a TMEM lane is an output row) which is what mode 1 consumes.
// mode 1 ("wgrad"):  dW[o, seg*Ki+i] = sum_{b,t} dZt[o,b,t] * Xt[i,b,t+shift[seg]]       weight-gradient
//                    both operands are the transposed activations (frames contiguous = K-major), reduction over all
//                    frames is split over `splits` CTAs per tile (fp32 partials, summed in fixed order afterwards).
#include "tc_common.cuh"
#include <stdlib.h>

namespace pka {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64, TC_STAGES = 3;   // 3 x 32 KB stages -> two CTAs per SM
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2, TC_B_BYTES = TC_BN * TC_BK * 2;
constexpr int TC_SMEM = TC_STAGES * (TC_A_BYTES + TC_B_BYTES) + 1024 /*align*/ + 256 /*barriers*/ + 512 /*bias*/;
constexpr int TC_THREADS = 192;

struct TcParams {
  int mode, Bt, T, N, K, nseg;
  int kb_per_seg, tiles_per_utt;
  int a_seg_col, b_seg_col;
  int shift[PKA_MAX_CTX];
  void* C; void* Ct;
  int ldc, c_dtype, Tp;
  const float* bias;
  int relu;
  int M;                       // mode 1: rows of dZt (output channels)
  int utt_per_split, tb_per_utt;
  int dbg;                     // diagnostics (env PKA_TC_DBG): 1 = skip the MMAs, 2 = skip the TMA loads
  pka_dropout drop;
  const __nv_bfloat16* addend; int ld_add;       // mode 0, optional: C = epi(acc) + addend[m, n]
};

// instruction descriptors: D=F32, A=B=BF16, N=128, M=128; both K-major (modes 0/1) or both MN-major (mode 2)
constexpr uint32_t kIdesc = make_idesc(TC_BM, TC_BN);
constexpr uint32_t kIdescMN = make_idesc(TC_BM, TC_BN, true, true);

// ---------------------------------------------------------------------------------------------- kernel
// bx / by / bz: the CTA's coordinates inside ITS problem's grid (== blockIdx for the single-problem kernel; decoded from
// a flat CTA index by the grouped kernel below)
__device__ __forceinline__ void gemm_tc_body(const CUtensorMap& mapA, const CUtensorMap& mapB, const TcParams& p,
                                             const int bx, const int by, const int bz) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + TC_STAGES * TC_A_BYTES;
  uint64_t* bars = (uint64_t*)(smem + TC_STAGES * (TC_A_BYTES + TC_B_BYTES));
  // bars[0..S) full, bars[S..2S) empty, bars[2S] accumulator ready; then the TMEM base address
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * TC_STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- tile coordinates
  int b0 = 0, t0 = 0, n0 = 0, seg_fixed = 0, num_k_iters, b_lo = 0;
  if (p.mode == 0) {
    b0 = bx / p.tiles_per_utt;
    t0 = (bx % p.tiles_per_utt) * TC_BM;
    n0 = by * TC_BN;
    num_k_iters = p.nseg * p.kb_per_seg;
  } else {
    t0 = bx * TC_BM;                               // output-channel (row) tile of dW
    const int i_tiles = (p.N + TC_BN - 1) / TC_BN;
    seg_fixed = by / i_tiles;
    n0 = (by % i_tiles) * TC_BN;
    b_lo = bz * p.utt_per_split;
    int b_hi = min(p.Bt, b_lo + p.utt_per_split);
    num_k_iters = max(0, b_hi - b_lo) * p.tb_per_utt;
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[TC_STAGES + s]), 1); }
    mbar_init(smem_u32(&bars[2 * TC_STAGES]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {                                 // 128 fp32 accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                    // the prologue above touched only shared / tensor memory

  if (warp == 0) {
    if (lane == 0) {                               // ===== TMA producer
      for (int i = 0; i < num_k_iters; ++i) {
        const int s = i % TC_STAGES, round = i / TC_STAGES;
        mbar_wait(smem_u32(&bars[TC_STAGES + s]), (round & 1) ^ 1);
        const uint32_t full = smem_u32(&bars[s]);
        if (p.dbg & 2) { mbar_expect_tx(full, 0); continue; }
        mbar_expect_tx(full, TC_A_BYTES + TC_B_BYTES);
        if (p.mode == 0) {
          const int seg = i / p.kb_per_seg, kb = i % p.kb_per_seg;
          tma_load_3d(smem_u32(sA + s * TC_A_BYTES), &mapA, full, seg * p.a_seg_col + kb * TC_BK, t0 + p.shift[seg], b0);
          tma_load_3d(smem_u32(sB + s * TC_B_BYTES), &mapB, full, seg * p.b_seg_col + kb * TC_BK, n0, 0);
        } else if (p.mode == 1) {
          const int b = b_lo + i / p.tb_per_utt, tb = i % p.tb_per_utt;
          tma_load_3d(smem_u32(sA + s * TC_A_BYTES), &mapA, full, tb * TC_BK, b, t0);
          tma_load_3d(smem_u32(sB + s * TC_B_BYTES), &mapB, full, tb * TC_BK + p.shift[seg_fixed], b, n0);
        } else {                                   // mode 2: [64 frames][64 channels] boxes, two per operand
          const int b = b_lo + i / p.tb_per_utt, tk = (i % p.tb_per_utt) * TC_BK;
          const uint32_t a = smem_u32(sA + s * TC_A_BYTES), bb = smem_u32(sB + s * TC_B_BYTES);
          tma_load_3d(a, &mapA, full, t0, tk, b);
          tma_load_3d(a + TC_A_BYTES / 2, &mapA, full, t0 + 64, tk, b);
          tma_load_3d(bb, &mapB, full, n0, tk + p.shift[seg_fixed], b);
          tma_load_3d(bb + TC_B_BYTES / 2, &mapB, full, n0 + 64, tk + p.shift[seg_fixed], b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                               // ===== MMA issuer
      for (int i = 0; i < num_k_iters; ++i) {
        const int s = i % TC_STAGES, round = i / TC_STAGES;
        mbar_wait(smem_u32(&bars[s]), round & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (p.dbg & 1) {
          if (i == 0) umma_f16(tmem_base, make_sdesc(smem_u32(sA)), make_sdesc(smem_u32(sB)), kIdesc, 0u);
        } else if (p.mode != 2) {
          const uint64_t da = make_sdesc(smem_u32(sA + s * TC_A_BYTES));
          const uint64_t db = make_sdesc(smem_u32(sB + s * TC_B_BYTES));
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)     // 16 bf16 = 32 B per UMMA_K step inside the 128 B swizzle atom
            umma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), kIdesc, (i | k) ? 1u : 0u);
        } else {
          const uint64_t da = make_sdesc(smem_u32(sA + s * TC_A_BYTES), TC_A_BYTES / 2);
          const uint64_t db = make_sdesc(smem_u32(sB + s * TC_B_BYTES), TC_B_BYTES / 2);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)     // 16 k-rows = two 8-row groups = 2048 B per UMMA_K step
            umma_f16(tmem_base, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), kIdescMN, (i | k) ? 1u : 0u);
        }
        umma_commit(smem_u32(&bars[TC_STAGES + s]));          // frees this smem stage when the MMAs have read it
      }
      umma_commit(smem_u32(&bars[2 * TC_STAGES]));            // accumulator complete
    }
  } else {                                         // ===== epilogue warps 2..5 -> TMEM lane quarter (warp % 4)
    const int q = warp & 3;
    const int row = q * 32 + lane;                 // row of the 128-row tile held by this thread
    // stage the bias of this CTA's 128 columns in shared memory while the main loop runs
    float* sbias = reinterpret_cast<float*>(smem + TC_STAGES * (TC_A_BYTES + TC_B_BYTES) + 256);   // 16-byte aligned
    {
      const int e = threadIdx.x - 64;              // 0..127
      const int n = n0 + e;
      sbias[e] = (p.mode == 0 && p.bias && n < p.N) ? p.bias[n] : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
    if (num_k_iters > 0) {
      mbar_wait(smem_u32(&bars[2 * TC_STAGES]), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const DropCtx dc = make_drop(p.drop);
    const float relu_floor = p.relu ? 0.f : -3.4e38f;
    const bool skip_all = (p.dbg & 8) != 0;
#pragma unroll 1
    for (int c = 0; c < (skip_all ? 0 : TC_BN / 32); ++c) {
      uint32_t r[32];
      if (num_k_iters > 0) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      const int nb = n0 + c * 32;
      if (nb >= p.N) break;                        // warp-uniform: nothing left in this tile
      const bool full = nb + 32 <= p.N;
      float v[32];
      if (p.mode == 0) {
        const int t = t0 + row;
        const bool valid = t < p.T && !(p.dbg & 4);
        const long long m = (long long)b0 * p.T + t;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bv = *reinterpret_cast<const float4*>(&sbias[c * 32 + j]);
          v[j] = fmaxf(__uint_as_float(r[j]) + bv.x, relu_floor);
          v[j + 1] = fmaxf(__uint_as_float(r[j + 1]) + bv.y, relu_floor);
          v[j + 2] = fmaxf(__uint_as_float(r[j + 2]) + bv.z, relu_floor);
          v[j + 3] = fmaxf(__uint_as_float(r[j + 3]) + bv.w, relu_floor);
        }
        if (dc.p > 0.f) {                          // N % 4 == 0 is required with dropout, so groups of 4 never straddle N
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (nb + j < p.N) {
              const float4 mul = dropout_mul4(dc, ((unsigned long long)m * p.N + nb + j) >> 2);
              v[j] *= mul.x; v[j + 1] *= mul.y; v[j + 2] *= mul.z; v[j + 3] *= mul.w;
            }
          }
        }
        if (p.addend && valid) {
          const __nv_bfloat16* ar = p.addend + m * p.ld_add + nb;
#pragma unroll
          for (int j = 0; j < 32; ++j) if (nb + j < p.N) v[j] += __bfloat162float(ar[j]);
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
          if (p.Ct) {                              // transposed copy Ct[n, b, t]: lanes = consecutive t -> coalesced
            __nv_bfloat16* dt = (__nv_bfloat16*)p.Ct + ((long long)b0 * p.Tp + t);
            const long long pitch = (long long)p.Bt * p.Tp;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (nb + j < p.N) dt[(long long)(nb + j) * pitch] = __float2bfloat16_rn(v[j]);
          }
        }
      } else {                                     // modes 1/2: fp32 partial of dW rows (output channels)
        const int o = t0 + row;
        if (o < p.M) {
          float* dst = (float*)p.C + ((long long)bz * p.M + o) * p.ldc + (long long)seg_fixed * p.N + nb;
          if (full && (p.ldc & 3) == 0 && (p.N & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                 __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (nb + j < p.N) dst[j] = __uint_as_float(r[j]);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
  }
}

__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcParams p) {
  gemm_tc_body(mapA, mapB, p, blockIdx.x, blockIdx.y, blockIdx.z);
}

// Grouped launch: up to TC_GROUP independent problems (the decoder-shaped weight gradients of one backward pass: 128- or
// 384-row outputs, reduction over the B*L tokens or B*T frames, ~32 CTAs and ~6 us each when launched alone) share ONE
// grid; a CTA finds its problem by a binary search over the prefix sums of the problems' grid sizes.  Tensor maps and
// parameters of all problems travel as __grid_constant__ kernel parameters (CUDA 12.1+: 32 KB parameter space).
constexpr int TC_GROUP = 40;
struct TcGroup {
  CUtensorMap mapA[TC_GROUP];
  CUtensorMap mapB[TC_GROUP];
  TcParams p[TC_GROUP];
  int cta_start[TC_GROUP + 1];
  int gx[TC_GROUP], gy[TC_GROUP];
  int n;
};
__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_group_kernel(const __grid_constant__ TcGroup g) {
  int lo = 0, hi = g.n;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (g.cta_start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid; }
  const int r = (int)blockIdx.x - g.cta_start[lo];
  const int gx = g.gx[lo], gxy = gx * g.gy[lo];
  gemm_tc_body(g.mapA[lo], g.mapB[lo], g.p[lo], r % gx, (r % gxy) / gx, r / gxy);
}

// fixed-order sum of the split partials: out[m, n] (+)= sum_s ws[s][m][n]
__global__ void tc_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out, long long per, int splits, int accumulate) {
  pdl_wait();
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < per; e += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += ws[(long long)sp * per + e];
    out[e] = accumulate ? out[e] + s : s;
  }
}

// fp32 master weight W[N, nseg*K] -> bf16 copy (forward operand) and bf16 Wd[K, nseg*N] with Wd[i, s*N+o] = W[o, s*K+i]
// (the data-gradient operand: K-major in the output-channel index)
__global__ void weight_relayout_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ Wf,
                                       __nv_bfloat16* __restrict__ Wd, int N, int K, int nseg) {
  pdl_wait();
  const long long total = (long long)N * nseg * K;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(e / ((long long)nseg * K));
    const int rem = (int)(e % ((long long)nseg * K));
    const int s = rem / K, i = rem % K;
    const __nv_bfloat16 v = __float2bfloat16_rn(W[e]);
    if (Wf) Wf[e] = v;
    if (Wd) Wd[(long long)i * nseg * N + (long long)s * N + o] = v;
  }
}

// per-head projection weights w_p[H, D, dk] (p < P, reference layout T/SubLayers.py:29-31) <-> the packed operands of the
// tensor-core GEMMs: Wf[(p*H+h)*dk + j, d] (forward, K-major in d) and Wd[d, (p*H+h)*dk + j] (data-gradient), bf16.
struct HeadPtrs { const float* w[3]; float* g[3]; };
__global__ void head_weight_relayout_kernel(HeadPtrs hp, __nv_bfloat16* __restrict__ Wf, __nv_bfloat16* __restrict__ Wd,
                                            int P, int H, int D, int dk) {
  pdl_wait();
  const long long per = (long long)H * D * dk, total = per * P;
  const int ntot = P * H * dk;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int pidx = (int)(e / per);
    const long long r = e % per;
    const int h = (int)(r / ((long long)D * dk)), d = (int)((r / dk) % D), j = (int)(r % dk);
    const __nv_bfloat16 v = __float2bfloat16_rn(hp.w[pidx][r]);
    const int n = (pidx * H + h) * dk + j;
    if (Wf) Wf[(long long)n * D + d] = v;
    if (Wd) Wd[(long long)d * ntot + n] = v;
  }
}
// dWcat fp32 [(p,h,j), d] -> dw_p[h, d, j]
__global__ void head_grad_relayout_kernel(HeadPtrs hp, const float* __restrict__ dWcat, int P, int H, int D, int dk) {
  pdl_wait();
  const long long per = (long long)H * D * dk, total = per * P;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int pidx = (int)(e / per);
    const long long r = e % per;
    const int h = (int)(r / ((long long)D * dk)), d = (int)((r / dk) % D), j = (int)(r % dk);
    if (hp.g[pidx]) hp.g[pidx][r] = dWcat[(long long)((pidx * H + h) * dk + j) * D + d];
  }
}

// dZ = (Y > 0) ? dY*scale : 0 in bf16, written row-major [rows, N] and transposed [N, Bt, Tp] (32x32 smem tiles)
template <typename Tin>
__global__ void relu_bwd_dual_kernel(const Tin* __restrict__ dY, const __nv_bfloat16* __restrict__ Y,
                                     __nv_bfloat16* __restrict__ dZ, __nv_bfloat16* __restrict__ dZt, int Bt, int T, int Tp,
                                     int N, float scale, int gate) {
  pdl_wait();
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, n = n0 + threadIdx.x;
    float v = 0.f;
    if (t < T && n < N) {
      const long long idx = ((long long)b * T + t) * N + n;
      v = to_f(dY[idx]);
      if (gate) v = __bfloat162float(Y[idx]) > 0.f ? v * scale : 0.f;
      if (dZ) dZ[idx] = __float2bfloat16_rn(v);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  if (dZt) {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int n = n0 + i, t = t0 + threadIdx.x;
      if (n < N && t < T) dZt[((long long)n * Bt + b) * Tp + t] = __float2bfloat16_rn(tile[threadIdx.x][i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------- host side
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// 3-D bf16 tensor map, innermost dim first; box = {64, box1, box2}; 128B swizzle; OOB elements read as zero
int make_map(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
             uint64_t stride2_bytes, uint32_t box1, uint32_t box2, const char* who, bool f32) {
  // cuTensorMapEncodeTiled is a driver call and needs a current context in THIS thread.  Autograd worker threads may
  // not have one bound yet when their first call into the library is a tensor-core GEMM, so make the (statically
  // linked) runtime bind the primary context first -- with a call that is legal during stream capture.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaStreamCaptureStatus cs;
    cudaStreamIsCapturing(cudaStreamPerThread, &cs);
    ctx_bound = true;
  }
  EncodeTiledFn enc = get_encode();
  PKA_REQUIRE(enc, PKA_EDEVICE, "%s: cuTensorMapEncodeTiled driver entry point not found", who);
  PKA_REQUIRE(aligned16(base) && stride1_bytes % 16 == 0 && stride2_bytes % 16 == 0, PKA_EALIGN,
              "%s: TMA needs 16-byte aligned base and strides (base %p, strides %llu, %llu)", who, base,
              (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes);
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {f32 ? 32u : 64u, box1, box2};              // 128 bytes along the contiguous dimension
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PKA_REQUIRE(r == CUDA_SUCCESS, PKA_EINVAL, "%s: cuTensorMapEncodeTiled failed with %d (dims %llu,%llu,%llu)", who, (int)r,
              (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2);
  return PKA_OK;
}

}  // namespace pka

namespace pka {
int launch_rows2(const pka_tc_desc* d, cudaStream_t st, bool* handled);    // gemm_tc_rows.cu
int launch_wgrad2(const pka_tc_desc* d, cudaStream_t st, bool* handled);   // gemm_tc_wgrad.cu
}

using namespace pka;

extern "C" int pka_gemm_tc(const pka_tc_desc* d, void* stream) {
  PKA_REQUIRE(d && d->A && d->B && d->C, PKA_EINVAL, "gemm_tc: null operand");
  PKA_REQUIRE(d->Bt > 0 && d->T > 0 && d->N > 0 && d->K > 0 && d->nseg >= 1 && d->nseg <= PKA_MAX_CTX, PKA_EINVAL,
              "gemm_tc: bad sizes Bt=%d T=%d N=%d K=%d nseg=%d", d->Bt, d->T, d->N, d->K, d->nseg);
  PKA_REQUIRE(!d->addend || (d->mode == 0 && d->ldadd >= d->N && !d->Ct), PKA_EINVAL, "gemm_tc: addend needs mode 0, ldadd >= N, no transposed copy");
  if (d->mode == 0) {                              // A-stationary / B-multicast kernel whenever the problem fits it
    PKA_REQUIRE(d->nseg == 1 || d->K % TC_BK == 0, PKA_EUNSUPPORTED, "gemm_tc: K=%d must be a multiple of %d when nseg>1", d->K, TC_BK);
    PKA_REQUIRE(d->drop.p == 0.f || d->N % 4 == 0, PKA_EUNSUPPORTED, "gemm_tc: dropout epilogue needs N%%4==0");
    bool handled = false;
    int rc2 = launch_rows2(d, as_stream(stream), &handled);
    if (rc2 || handled) return rc2;
  }
  if (d->mode == 2 && d->M > 0 && d->splits >= 1) {   // CTA-pair weight-gradient kernel whenever the tile shape allows
    bool handled = false;
    int rc2 = launch_wgrad2(d, as_stream(stream), &handled);
    if (rc2 || handled) return rc2;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    PKA_REQUIRE(e == cudaSuccess, PKA_ELAUNCH, "gemm_tc: cannot opt in to %d bytes of shared memory: %s", TC_SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  TcParams p;
  p.mode = d->mode; p.Bt = d->Bt; p.T = d->T; p.N = d->N; p.K = d->K; p.nseg = d->nseg;
  p.a_seg_col = d->a_seg_col; p.b_seg_col = d->b_seg_col;
  for (int i = 0; i < PKA_MAX_CTX; ++i) p.shift[i] = d->shift[i];
  p.C = d->C; p.Ct = d->Ct; p.ldc = d->ldc; p.c_dtype = d->c_dtype; p.Tp = d->Tp;
  p.bias = d->bias; p.relu = d->relu; p.M = d->M; p.drop = d->drop;
  p.addend = d->mode == 0 ? (const __nv_bfloat16*)d->addend : nullptr; p.ld_add = d->ldadd;
  p.kb_per_seg = (d->K + TC_BK - 1) / TC_BK;
  p.tiles_per_utt = (d->T + TC_BM - 1) / TC_BM;
  p.utt_per_split = 0; p.tb_per_utt = 0;
  { const char* e = getenv("PKA_TC_DBG"); p.dbg = e ? atoi(e) : 0; }
  CUtensorMap mapA, mapB;
  dim3 grid;
  int rc;
  if (d->mode == 0) {
    PKA_REQUIRE(d->nseg == 1 || d->K % TC_BK == 0, PKA_EUNSUPPORTED, "gemm_tc: K=%d must be a multiple of %d when nseg>1", d->K, TC_BK);
    PKA_REQUIRE(d->drop.p == 0.f || d->N % 4 == 0, PKA_EUNSUPPORTED, "gemm_tc: dropout epilogue needs N%%4==0");
    PKA_REQUIRE(!d->Ct || d->Tp >= d->T, PKA_EINVAL, "gemm_tc: Tp < T");
    // A: [Bt, T, lda] bf16, visible width = a_seg_col*(nseg-1)+K columns; B: weights [N, ldb] with nseg K-segments
    const uint64_t a_cols = (uint64_t)d->a_seg_col * (d->nseg - 1) + d->K;
    rc = make_map(&mapA, d->A, a_cols, d->T, d->Bt, (uint64_t)d->lda * 2, (uint64_t)d->T * d->lda * 2, TC_BM, 1, "gemm_tc A");
    if (rc) return rc;
    const uint64_t b_cols = (uint64_t)d->b_seg_col * (d->nseg - 1) + d->K;
    rc = make_map(&mapB, d->B, b_cols, d->N, 1, (uint64_t)d->ldb * 2, (uint64_t)d->N * d->ldb * 2, TC_BN, 1, "gemm_tc B");
    if (rc) return rc;
    grid = dim3(p.tiles_per_utt * d->Bt, (d->N + TC_BN - 1) / TC_BN, 1);
  } else if (d->mode == 2) {
    PKA_REQUIRE(d->M > 0 && d->splits >= 1, PKA_EINVAL, "gemm_tc wgrad: M=%d splits=%d", d->M, d->splits);
    PKA_REQUIRE(d->c_dtype == PKA_F32, PKA_EUNSUPPORTED, "gemm_tc wgrad: partial sums are fp32");
    // A = dZ [Bt, T, lda] (M = output channels contiguous), B = X [Bt, T, ldb] (N = input channels contiguous):
    // MN-major operands, the reduction runs over frames; boxes are [64 frames][64 channels]
    rc = make_map(&mapA, d->A, d->M, d->T, d->Bt, (uint64_t)d->lda * 2, (uint64_t)d->T * d->lda * 2, 64, 1, "gemm_tc dZ");
    if (rc) return rc;
    rc = make_map(&mapB, d->B, d->N, d->T, d->Bt, (uint64_t)d->ldb * 2, (uint64_t)d->T * d->ldb * 2, 64, 1, "gemm_tc X");
    if (rc) return rc;
    p.utt_per_split = (d->Bt + d->splits - 1) / d->splits;
    p.tb_per_utt = (d->T + TC_BK - 1) / TC_BK;
    grid = dim3((d->M + TC_BM - 1) / TC_BM, ((d->N + TC_BN - 1) / TC_BN) * d->nseg, d->splits);
  } else {
    PKA_REQUIRE(d->M > 0 && d->splits >= 1 && d->Tp >= d->T && d->Tp % 8 == 0, PKA_EINVAL, "gemm_tc wgrad: M=%d splits=%d Tp=%d", d->M, d->splits, d->Tp);
    PKA_REQUIRE(d->c_dtype == PKA_F32, PKA_EUNSUPPORTED, "gemm_tc wgrad: partial sums are fp32");
    // A = dZt [M, Bt, Tp], B = Xt [N, Bt, Tp]: dims {T, Bt, rows}, the box covers 64 frames of one utterance, 128 rows
    rc = make_map(&mapA, d->A, d->T, d->Bt, d->M, (uint64_t)d->Tp * 2, (uint64_t)d->Bt * d->Tp * 2, 1, TC_BM, "gemm_tc dZt");
    if (rc) return rc;
    rc = make_map(&mapB, d->B, d->T, d->Bt, d->N, (uint64_t)d->Tp * 2, (uint64_t)d->Bt * d->Tp * 2, 1, TC_BN, "gemm_tc Xt");
    if (rc) return rc;
    p.utt_per_split = (d->Bt + d->splits - 1) / d->splits;
    p.tb_per_utt = (d->T + TC_BK - 1) / TC_BK;
    grid = dim3((d->M + TC_BM - 1) / TC_BM, ((d->N + TC_BN - 1) / TC_BN) * d->nseg, d->splits);
  }
  launch_k(gemm_tc_kernel, grid, TC_THREADS, TC_SMEM, as_stream(stream), mapA, mapB, p);
  return check_launch("gemm_tc");
}

// Weight gradients (mode 2) of several independent problems in one launch (see gemm_tc_group_kernel).  Every descriptor
// is what pka_gemm_tc would take; problems the CTA-pair kernel would handle are accepted too (they run on this kernel).
extern "C" int pka_gemm_tc_wgrad_group(const pka_tc_desc* descs, int n_desc, void* stream) {
  PKA_REQUIRE(descs && n_desc > 0, PKA_EINVAL, "gemm_tc_wgrad_group: empty table");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    PKA_REQUIRE(e == cudaSuccess, PKA_ELAUNCH, "gemm_tc_wgrad_group: cannot opt in to %d bytes of shared memory: %s", TC_SMEM, cudaGetErrorString(e));
    attr_set = true;
  }
  int dbg = 0;
  { const char* e = getenv("PKA_TC_DBG"); dbg = e ? atoi(e) : 0; }
  for (int base = 0; base < n_desc; base += TC_GROUP) {
    static thread_local TcGroup g;                 // 20 KB: keep it off the stack
    const int n = n_desc - base < TC_GROUP ? n_desc - base : TC_GROUP;
    int ctas = 0;
    for (int i = 0; i < n; ++i) {
      const pka_tc_desc* d = descs + base + i;
      PKA_REQUIRE(d->A && d->B && d->C && d->mode == 2, PKA_EINVAL, "gemm_tc_wgrad_group: problem %d: mode-2 descriptor with operands expected", base + i);
      PKA_REQUIRE(d->Bt > 0 && d->T > 0 && d->M > 0 && d->N > 0 && d->nseg >= 1 && d->nseg <= PKA_MAX_CTX && d->splits >= 1 && d->c_dtype == PKA_F32,
                  PKA_EINVAL, "gemm_tc_wgrad_group: problem %d: bad sizes", base + i);
      TcParams& p = g.p[i];
      p.mode = 2; p.Bt = d->Bt; p.T = d->T; p.N = d->N; p.K = d->K; p.nseg = d->nseg;
      p.a_seg_col = d->a_seg_col; p.b_seg_col = d->b_seg_col;
      for (int k = 0; k < PKA_MAX_CTX; ++k) p.shift[k] = d->shift[k];
      p.C = d->C; p.Ct = nullptr; p.ldc = d->ldc; p.c_dtype = d->c_dtype; p.Tp = 0;
      p.bias = nullptr; p.relu = 0; p.M = d->M; p.drop = no_dropout(); p.addend = nullptr; p.ld_add = 0;
      p.kb_per_seg = (d->K + TC_BK - 1) / TC_BK;
      p.tiles_per_utt = (d->T + TC_BM - 1) / TC_BM;
      p.utt_per_split = (d->Bt + d->splits - 1) / d->splits;
      p.tb_per_utt = (d->T + TC_BK - 1) / TC_BK;
      p.dbg = dbg;
      int rc = make_map(&g.mapA[i], d->A, d->M, d->T, d->Bt, (uint64_t)d->lda * 2, (uint64_t)d->T * d->lda * 2, 64, 1, "gemm_tc_wgrad_group dZ");
      if (rc) return rc;
      rc = make_map(&g.mapB[i], d->B, d->N, d->T, d->Bt, (uint64_t)d->ldb * 2, (uint64_t)d->T * d->ldb * 2, 64, 1, "gemm_tc_wgrad_group X");
      if (rc) return rc;
      g.gx[i] = (d->M + TC_BM - 1) / TC_BM;
      g.gy[i] = ((d->N + TC_BN - 1) / TC_BN) * d->nseg;
      g.cta_start[i] = ctas;
      ctas += g.gx[i] * g.gy[i] * d->splits;
    }
    g.cta_start[n] = ctas;
    g.n = n;
    launch_k(gemm_tc_group_kernel, ctas, TC_THREADS, TC_SMEM, as_stream(stream), g);
    int rc = check_launch("gemm_tc_wgrad_group");
    if (rc) return rc;
  }
  return PKA_OK;
}

extern "C" int pka_tc_reduce(const float* ws, float* out, int64_t per, int splits, int accumulate, void* stream) {
  PKA_REQUIRE(ws && out && per > 0 && splits >= 1, PKA_EINVAL, "tc_reduce: bad arguments");
  long long blocks = (per + 255) / 256;
  if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
  launch_k(tc_reduce_kernel, (int)blocks, 256, 0, as_stream(stream), ws, out, per, splits, accumulate);
  return check_launch("tc_reduce");
}

extern "C" int pka_weight_relayout(const float* W, void* Wf, void* Wd, int N, int K, int nseg, void* stream) {
  PKA_REQUIRE(W && (Wf || Wd) && N > 0 && K > 0 && nseg >= 1, PKA_EINVAL, "weight_relayout: bad arguments");
  const long long total = (long long)N * nseg * K;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
  launch_k(weight_relayout_kernel, (int)blocks, 256, 0, as_stream(stream), W, (__nv_bfloat16*)Wf, (__nv_bfloat16*)Wd, N, K, nseg);
  return check_launch("weight_relayout");
}

extern "C" int pka_relu_bwd_dual(const void* dY, int dy_dtype, const void* Y, void* dZ, void* dZt, int Bt, int T, int Tp, int N,
                                 float scale, int gate, void* stream) {
  PKA_REQUIRE(dY && (dZ || dZt) && (!gate || Y), PKA_EINVAL, "relu_bwd_dual: null pointer");
  PKA_REQUIRE(Bt > 0 && Bt <= 65535 && T > 0 && N > 0 && Tp >= T, PKA_EINVAL, "relu_bwd_dual: bad sizes");
  dim3 grid((N + 31) / 32, (T + 31) / 32, Bt), block(32, 8);
  if (dy_dtype == PKA_BF16)
    launch_k(relu_bwd_dual_kernel<__nv_bfloat16>, grid, block, 0, as_stream(stream), (const __nv_bfloat16*)dY, (const __nv_bfloat16*)Y, (__nv_bfloat16*)dZ, (__nv_bfloat16*)dZt, Bt, T, Tp, N, scale, gate);
  else if (dy_dtype == PKA_F32)
    launch_k(relu_bwd_dual_kernel<float>, grid, block, 0, as_stream(stream), (const float*)dY, (const __nv_bfloat16*)Y, (__nv_bfloat16*)dZ, (__nv_bfloat16*)dZt, Bt, T, Tp, N, scale, gate);
  else PKA_REQUIRE(false, PKA_EUNSUPPORTED, "relu_bwd_dual: dtype %d", dy_dtype);
  return check_launch("relu_bwd_dual");
}

extern "C" int pka_head_weight_relayout(const float* w0, const float* w1, const float* w2, int P, int H, int D, int dk,
                                        void* Wf, void* Wd, void* stream) {
  PKA_REQUIRE(w0 && P >= 1 && P <= 3 && (P < 2 || w1) && (P < 3 || w2) && (Wf || Wd), PKA_EINVAL, "head_weight_relayout: bad arguments");
  HeadPtrs hp; hp.w[0] = w0; hp.w[1] = w1; hp.w[2] = w2; hp.g[0] = hp.g[1] = hp.g[2] = nullptr;
  const long long total = (long long)P * H * D * dk;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
  launch_k(head_weight_relayout_kernel, (int)blocks, 256, 0, as_stream(stream), hp, (__nv_bfloat16*)Wf, (__nv_bfloat16*)Wd, P, H, D, dk);
  return check_launch("head_weight_relayout");
}

// fixed-order sum of the split partials of a packed head-projection weight gradient, written straight in the
// reference's per-head layout: g_p[h, d, j] = sum_s ws[s][(p*H + h)*dk + j][d]   (tc_reduce + head_grad_relayout in one)
__global__ void tc_reduce_heads_kernel(HeadPtrs hp, const float* __restrict__ ws, int splits, int P, int H, int D, int dk) {
  pdl_wait();
  const long long per = (long long)P * H * dk * D;            // elements of one partial [(p,h,j), d]
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < per; e += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(e % D);
    const int n = (int)(e / D);
    const int j = n % dk, h = (n / dk) % H, pidx = n / (dk * H);
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += ws[(long long)sp * per + e];
    if (hp.g[pidx]) hp.g[pidx][((long long)h * D + d) * dk + j] = s;
  }
}

extern "C" int pka_tc_reduce_heads(const float* ws, float* g0, float* g1, float* g2, int splits, int P, int H, int D, int dk,
                                   void* stream) {
  PKA_REQUIRE(ws && splits >= 1 && P >= 1 && P <= 3 && H > 0 && D > 0 && dk > 0, PKA_EINVAL, "tc_reduce_heads: bad arguments");
  HeadPtrs hp; hp.w[0] = hp.w[1] = hp.w[2] = nullptr; hp.g[0] = g0; hp.g[1] = g1; hp.g[2] = g2;
  const long long total = (long long)P * H * D * dk;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
  launch_k(tc_reduce_heads_kernel, (int)blocks, 256, 0, as_stream(stream), hp, ws, splits, P, H, D, dk);
  return check_launch("tc_reduce_heads");
}

extern "C" int pka_head_grad_relayout(const float* dWcat, float* g0, float* g1, float* g2, int P, int H, int D, int dk,
                                      void* stream) {
  PKA_REQUIRE(dWcat && P >= 1 && P <= 3, PKA_EINVAL, "head_grad_relayout: bad arguments");
  HeadPtrs hp; hp.w[0] = hp.w[1] = hp.w[2] = nullptr; hp.g[0] = g0; hp.g[1] = g1; hp.g[2] = g2;
  const long long total = (long long)P * H * D * dk;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
  launch_k(head_grad_relayout_kernel, (int)blocks, 256, 0, as_stream(stream), hp, dWcat, P, H, D, dk);
  return check_launch("head_grad_relayout");
}
