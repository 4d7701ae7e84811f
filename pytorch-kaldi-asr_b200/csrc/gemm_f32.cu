// fp32 SIMT GEMM family -- the "exact" path (bit-stable fp32 FMA accumulation in a fixed order), used for fp32
// parity against the CPU oracle and for beam search, where token sequences must not flip.  The bf16 tensor-core
// path lives in gemm_tc.cu.  Contract: include/pka_b200.h (pka_gemm_desc).
//
// One kernel template covers forward, data-gradient and weight-gradient of every linear map on the path
// (BottleLinear, TDNN splice+Linear, per-head projections, Conv1d k=1) through four operand layouts, K "segments"
// (splice contexts / heads) and a batch axis.  Frame splicing (reference ConcatLayer, L/pytorch/TDNN.py:20-28) is a
// row shift with a bounds predicate inside the tile loader, so the [B,T,n_ctx*D] tensor is never materialised.
#include "common.cuh"
#include <stdlib.h>

namespace pka {

struct GemmParams {
  const float* A; const float* B; float* C;
  const float* bias; const float* residual;
  int M, N, K, nseg;
  int lda, ldb, ldc, ldr;
  long long a_seg_off, b_seg_off, a_batch_off, b_batch_off, c_batch_off;
  int shiftA[PKA_MAX_CTX];
  int shiftB[PKA_MAX_CTX];
  int T;
  int relu, accumulate;
  int splitk, nbatch;          // split-K: grid.z = nbatch*splitk, raw partial sums go to ws[split][batch][M][N]
  float* ws;
  pka_dropout drop;
};

template <int BM, int BN, int BK, int TM, int TN, bool TA, bool TB>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_f32_kernel(const GemmParams p) {
  pdl_wait();
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
  static_assert(TM % 4 == 0 && TN % 4 == 0, "micro-tile is built from float4 groups");
  static_assert(NT % 32 == 0 && BK == 32, "loader assumes one warp spans the contiguous tile dimension");
  __shared__ __align__(16) float As[BK][LDA_S];
  __shared__ __align__(16) float Bs[BK][LDB_S];

  const int tid = threadIdx.x, lane = tid & 31, wrow = tid >> 5;
  constexpr int NW = NT / 32;
  const int m_blk = blockIdx.y * BM, n_blk = blockIdx.x * BN;
  const int batch = blockIdx.z / p.splitk, split = blockIdx.z % p.splitk;
  // split-K (nseg == 1 only): this CTA reduces k in [k_begin, k_end)
  const int k_chunk = ((p.K + p.splitk - 1) / p.splitk + BK - 1) / BK * BK;
  const int k_begin = split * k_chunk;
  const int k_end = min(p.K, k_begin + k_chunk);
  const float* __restrict__ Ab = p.A + (long long)batch * p.a_batch_off;
  const float* __restrict__ Bb = p.B + (long long)batch * p.b_batch_off;
  float* __restrict__ Cb = p.C + (long long)batch * p.c_batch_off;
  const int shiftB = p.shiftB[batch < PKA_MAX_CTX ? batch : 0];

  // micro-tile: TM rows as TM/4 groups of 4 spaced BM/(TM/4) apart (conflict-free float4 smem reads), same for cols
  constexpr int GM = TM / 4, GN = TN / 4;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // ---- tile loaders: every thread first issues ALL its global loads of a tile into registers (compile-time trip
  // counts -> the loads are in flight together), the registers are stored to shared memory after the previous tile's
  // math, and the next tile's loads are issued before the math of the current one (register double buffering).
  constexpr int RA = TA ? (BK / NW) * (BM / 32) : BM / NW;
  constexpr int RB = TB ? BN / NW : (BK / NW) * (BN / 32);
  float ra[RA], rb[RB];
  const int k_iters = (k_end - k_begin + BK - 1) / BK;
  const int n_it = k_end > k_begin ? p.nseg * k_iters : 0;

  auto load_tile = [&](int it) {
    const int seg = it / k_iters, k0 = k_begin + (it % k_iters) * BK;
    const float* __restrict__ As_g = Ab + (long long)seg * p.a_seg_off;
    const float* __restrict__ Bs_g = Bb + (long long)seg * p.b_seg_off;
    if (!TA) {                        // A[M,K] row-major: lanes along k
      const int shA = p.shiftA[seg < PKA_MAX_CTX ? seg : 0];
      const int k = k0 + lane;
#pragma unroll
      for (int i = 0; i < RA; ++i) {
        const int m = m_blk + wrow + i * NW;
        float v = 0.f;
        if (m < p.M && k < k_end) {
          bool ok = true;
          if (p.T > 0) { int t = m % p.T + shA; ok = (t >= 0) && (t < p.T); }
          if (ok) v = As_g[(long long)(m + shA) * p.lda + k];
        }
        ra[i] = v;
      }
    } else {                          // A stored [K,M]: lanes along m
#pragma unroll
      for (int i = 0; i < BK / NW; ++i) {
        const int k = k0 + wrow + i * NW;
#pragma unroll
        for (int c = 0; c < BM / 32; ++c) {
          const int m = m_blk + lane + 32 * c;
          ra[i * (BM / 32) + c] = (k < k_end && m < p.M) ? As_g[(long long)k * p.lda + m] : 0.f;
        }
      }
    }
    if (TB) {                         // B stored [N,K] (nn.Linear weight): lanes along k
      const int k = k0 + lane;
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int n = n_blk + wrow + i * NW;
        rb[i] = (n < p.N && k < k_end) ? Bs_g[(long long)n * p.ldb + k] : 0.f;
      }
    } else {                          // B stored [K,N]: lanes along n; optional frame shift on the reduction index
#pragma unroll
      for (int i = 0; i < BK / NW; ++i) {
        const int k = k0 + wrow + i * NW;
        bool ok = k < k_end;
        if (ok && p.T > 0) { int t = k % p.T + shiftB; ok = (t >= 0) && (t < p.T); }
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          const int n = n_blk + lane + 32 * c;
          rb[i * (BN / 32) + c] = (ok && n < p.N) ? Bs_g[(long long)(k + shiftB) * p.ldb + n] : 0.f;
        }
      }
    }
  };
  auto store_tile = [&]() {
    if (!TA) {
#pragma unroll
      for (int i = 0; i < RA; ++i) As[lane][wrow + i * NW] = ra[i];
    } else {
#pragma unroll
      for (int i = 0; i < BK / NW; ++i)
#pragma unroll
        for (int c = 0; c < BM / 32; ++c) As[wrow + i * NW][lane + 32 * c] = ra[i * (BM / 32) + c];
    }
    if (TB) {
#pragma unroll
      for (int i = 0; i < RB; ++i) Bs[lane][wrow + i * NW] = rb[i];
    } else {
#pragma unroll
      for (int i = 0; i < BK / NW; ++i)
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) Bs[wrow + i * NW][lane + 32 * c] = rb[i * (BN / 32) + c];
    }
  };

  if (n_it > 0) load_tile(0);
  for (int it = 0; it < n_it; ++it) {
    store_tile();
    __syncthreads();
    if (it + 1 < n_it) load_tile(it + 1);
#pragma unroll 8
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < GM; ++g) {
        float4 v = *reinterpret_cast<const float4*>(&As[kk][g * (BM / GM) + ty * 4]);
        a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < GN; ++g) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[kk][g * (BN / GN) + tx * 4]);
        b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  if (p.splitk > 1) {              // raw partial sums; splitk_reduce_kernel finishes in a fixed order
    float* wsb = p.ws + ((long long)split * p.nbatch + batch) * p.M * p.N;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m_blk + (i / 4) * (BM / GM) + ty * 4 + (i % 4);
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n_blk + (j / 4) * (BN / GN) + tx * 4 + (j % 4);
        if (n < p.N) wsb[(long long)m * p.N + n] = acc[i][j];
      }
    }
    return;
  }
  // ---- epilogue: bias -> ReLU -> dropout -> + residual -> (accumulate) -> store
  DropCtx dc = make_drop(p.drop);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m_blk + (i / 4) * (BM / GM) + ty * 4 + (i % 4);
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n_blk + (j / 4) * (BN / GN) + tx * 4 + (j % 4);
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (p.relu) v = fmaxf(v, 0.f);
      if (dc.p > 0.f) {
        unsigned long long idx = ((unsigned long long)batch * p.M + m) * (unsigned long long)p.N + n;
        v = dropout_keep(dc, idx) ? v * dc.scale : 0.f;
      }
      if (p.residual) v += p.residual[(long long)m * p.ldr + n];
      float* dst = Cb + (long long)m * p.ldc + n;
      if (p.accumulate) v += *dst;
      *dst = v;
    }
  }
}

// Small-M variant for the decoder-step GEMMs (M = live hypotheses, K <= 256): the whole reduction dimension of a
// [32 rows] x [64 columns] output tile is staged in shared memory with ONE round of 16-byte loads (the generic kernel
// above pays a global-memory latency per 32-wide K tile, which is all these launches consist of), then each thread owns
// 4 rows x 2 columns.  Same accumulation order as the generic kernel (one accumulator per output, k ascending).
constexpr int SM_BM = 32, SM_BN = 64;
__global__ void __launch_bounds__(256)
gemm_f32_small_kernel(const GemmParams p) {
  pdl_wait();
  extern __shared__ __align__(16) float smf[];
  const int K = p.K, ldb_s = K + 4;
  float* As = smf;                                 // [32][K]
  float* Bs = smf + SM_BM * K;                     // [64][K + 4]
  const int m_blk = blockIdx.y * SM_BM, n_blk = blockIdx.x * SM_BN;
  const int k4n = K >> 2;
  for (int e = threadIdx.x; e < SM_BM * k4n; e += 256) {
    const int r = e / k4n, c = e - r * k4n, m = m_blk + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < p.M) v = *reinterpret_cast<const float4*>(p.A + (long long)m * p.lda + c * 4);
    *reinterpret_cast<float4*>(As + r * K + c * 4) = v;
  }
  for (int e = threadIdx.x; e < SM_BN * k4n; e += 256) {
    const int r = e / k4n, c = e - r * k4n, n = n_blk + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < p.N) v = *reinterpret_cast<const float4*>(p.B + (long long)n * p.ldb + c * 4);
    *reinterpret_cast<float4*>(Bs + r * ldb_s + c * 4) = v;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
  const float* a0 = As + (ty * 4) * K;
  const float* b0 = Bs + tx * ldb_s;
  const float* b1 = Bs + (tx + 32) * ldb_s;
#pragma unroll 4
  for (int c = 0; c < k4n; ++c) {
    const float4 bv0 = *reinterpret_cast<const float4*>(b0 + c * 4);
    const float4 bv1 = *reinterpret_cast<const float4*>(b1 + c * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 av = *reinterpret_cast<const float4*>(a0 + i * K + c * 4);
      acc[i][0] = fmaf(av.x, bv0.x, acc[i][0]); acc[i][0] = fmaf(av.y, bv0.y, acc[i][0]);
      acc[i][0] = fmaf(av.z, bv0.z, acc[i][0]); acc[i][0] = fmaf(av.w, bv0.w, acc[i][0]);
      acc[i][1] = fmaf(av.x, bv1.x, acc[i][1]); acc[i][1] = fmaf(av.y, bv1.y, acc[i][1]);
      acc[i][1] = fmaf(av.z, bv1.z, acc[i][1]); acc[i][1] = fmaf(av.w, bv1.w, acc[i][1]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m_blk + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n_blk + tx + 32 * j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.residual) v += p.residual[(long long)m * p.ldr + n];
      float* dst = p.C + (long long)m * p.ldc + n;
      if (p.accumulate) v += *dst;
      *dst = v;
    }
  }
}

__global__ void splitk_reduce_kernel(const GemmParams p) {
  pdl_wait();
  const long long per = (long long)p.M * p.N, total = per * p.nbatch;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int batch = (int)(e / per);
    const long long r = e % per;
    const int m = (int)(r / p.N), n = (int)(r % p.N);
    float s = 0.f;
    for (int sp = 0; sp < p.splitk; ++sp) s += p.ws[(long long)sp * total + e];
    float* dst = p.C + (long long)batch * p.c_batch_off + (long long)m * p.ldc + n;
    *dst = p.accumulate ? *dst + s : s;
  }
}

template <int BM, int BN, int TM, int TN>
static int launch_cfg(const GemmParams& p, int transA, int transB, int nbatch, cudaStream_t st) {
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, nbatch * p.splitk);
  dim3 block((BM / TM) * (BN / TN));
  if (!transA && transB) launch_k(gemm_f32_kernel<BM, BN, 32, TM, TN, false, true>, grid, block, 0, st, p);
  else if (!transA && !transB) launch_k(gemm_f32_kernel<BM, BN, 32, TM, TN, false, false>, grid, block, 0, st, p);
  else if (transA && !transB) launch_k(gemm_f32_kernel<BM, BN, 32, TM, TN, true, false>, grid, block, 0, st, p);
  else launch_k(gemm_f32_kernel<BM, BN, 32, TM, TN, true, true>, grid, block, 0, st, p);
  return check_launch("gemm_f32");
}

}  // namespace pka

extern "C" int pka_gemm_f32(const pka_gemm_desc* d, void* stream) {
  using namespace pka;
  PKA_REQUIRE(d && d->A && d->B && d->C, PKA_EINVAL, "gemm_f32: null operand");
  PKA_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0 && d->nseg >= 1 && d->nbatch >= 1, PKA_EINVAL,
              "gemm_f32: bad sizes M=%d N=%d K=%d nseg=%d nbatch=%d", d->M, d->N, d->K, d->nseg, d->nbatch);
  PKA_REQUIRE(d->nseg <= PKA_MAX_CTX || d->T == 0, PKA_EUNSUPPORTED, "gemm_f32: more than %d spliced segments", PKA_MAX_CTX);
  PKA_REQUIRE(d->nbatch <= PKA_MAX_CTX || d->T == 0 || d->transB, PKA_EUNSUPPORTED, "gemm_f32: more than %d shifted batches", PKA_MAX_CTX);
  PKA_REQUIRE(d->T == 0 || (d->transA ? true : d->M % d->T == 0), PKA_EINVAL, "gemm_f32: M=%d is not a multiple of T=%d", d->M, d->T);
  PKA_REQUIRE(d->T == 0 || !d->transA || d->transB || d->K % d->T == 0, PKA_EINVAL, "gemm_f32: K=%d is not a multiple of T=%d", d->K, d->T);
  PKA_REQUIRE(d->nbatch <= 65535, PKA_EUNSUPPORTED, "gemm_f32: nbatch too large");
  GemmParams p;
  p.A = (const float*)d->A; p.B = (const float*)d->B; p.C = (float*)d->C;
  p.bias = d->bias; p.residual = (const float*)d->residual;
  p.M = d->M; p.N = d->N; p.K = d->K; p.nseg = d->nseg;
  p.lda = d->lda; p.ldb = d->ldb; p.ldc = d->ldc; p.ldr = d->ldr;
  p.a_seg_off = d->a_seg_off; p.b_seg_off = d->b_seg_off;
  p.a_batch_off = d->a_batch_off; p.b_batch_off = d->b_batch_off; p.c_batch_off = d->c_batch_off;
  for (int i = 0; i < PKA_MAX_CTX; ++i) { p.shiftA[i] = d->T > 0 ? d->shiftA[i] : 0; p.shiftB[i] = d->T > 0 ? d->shiftB[i] : 0; }
  p.T = d->T; p.relu = d->relu; p.accumulate = d->accumulate; p.drop = d->drop;
  p.splitk = d->splitk > 1 ? d->splitk : 1; p.nbatch = d->nbatch; p.ws = (float*)d->splitk_ws;
  if (p.splitk > 1) {
    PKA_REQUIRE(d->nseg == 1 && !d->bias && !d->relu && !d->residual && d->drop.p == 0.f, PKA_EUNSUPPORTED, "gemm_f32: split-K needs nseg=1 and no epilogue");
    PKA_REQUIRE(p.ws, PKA_EINVAL, "gemm_f32: split-K workspace missing");
    PKA_REQUIRE((long long)d->nbatch * p.splitk <= 65535, PKA_EUNSUPPORTED, "gemm_f32: nbatch*splitk too large");
  }
  cudaStream_t st = as_stream(stream);
  // decoder-step shape: few rows, short reduction, nn.Linear weight layout, no splice / dropout / batching
  if (!d->transA && d->transB && d->nseg == 1 && d->nbatch == 1 && d->T == 0 && p.splitk == 1 && d->drop.p == 0.f &&
      d->M <= 4096 && d->K <= 256 && d->K % 4 == 0 && d->lda % 4 == 0 && d->ldb % 4 == 0 && aligned16(d->A) && aligned16(d->B) &&
      !getenv("PKA_GEMM_NOSMALL")) {
    const int smem = (SM_BM * d->K + SM_BN * (d->K + 4)) * (int)sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(gemm_f32_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      PKA_REQUIRE(e == cudaSuccess, PKA_ELAUNCH, "gemm_f32: cannot opt in to shared memory: %s", cudaGetErrorString(e));
      attr_set = true;
    }
    launch_k(gemm_f32_small_kernel, dim3((d->N + SM_BN - 1) / SM_BN, (d->M + SM_BM - 1) / SM_BM), 256, (size_t)smem, st, p);
    return check_launch("gemm_f32(small)");
  }
  // big tiles only when they still give >= ~1 wave of CTAs
  long long big_ctas = (long long)((d->M + 127) / 128) * ((d->N + 127) / 128) * d->nbatch * p.splitk;
  bool big = d->N >= 128 && big_ctas >= kNumSMs;
  if (const char* force = getenv("PKA_GEMM_TILE")) big = (force[0] == 'b');     // diagnostics only
  int rc = big ? launch_cfg<128, 128, 8, 8>(p, d->transA, d->transB, d->nbatch, st)
               : launch_cfg<64, 64, 4, 4>(p, d->transA, d->transB, d->nbatch, st);
  if (rc || p.splitk == 1) return rc;
  long long total = (long long)p.M * p.N * p.nbatch;
  int blocks = (int)((total + 255) / 256 < (long long)kNumSMs * 8 ? (total + 255) / 256 : (long long)kNumSMs * 8);
  launch_k(splitk_reduce_kernel, blocks, 256, 0, st, p);
  return check_launch("gemm_f32 split-K reduce");
}
