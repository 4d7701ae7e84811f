// Batched helper launches of the training step: ONE kernel finishes every split reduction a backward pass has left
// behind, ONE kernel refreshes every bf16 GEMM operand copy of the fp32 master weights.
//
// Why: the TIMIT-config step is launch/latency bound.  Round 1 ran 57 "finish" kernels
// per step (fixed-order sums of weight-gradient split partials, bias-gradient column sums, LayerNorm gain/offset
// gradients, per-head gradient relayouts: ~190 us) and 29 operand relayouts (~86 us), each a 2-4 us launch moving a few
// KB.  None of them is on the critical path of backward: the partials only have to be summed before the optimiser (or
// the gradient all-reduce) reads the gradient arena, and the operand copies only change when Adam changes the weights.
// Both become a table of jobs executed by one grid, with the same fixed summation order as before (bit-reproducible).
#include "common.cuh"

namespace pka {

// ---------------------------------------------------------------------------------------------- reduction jobs
struct ReduceTable {
  pka_reduce_job job[PKA_MAX_REDUCE_JOBS];
  int32_t cta_start[PKA_MAX_REDUCE_JOBS + 1];      // first CTA of every job (prefix sums of ceil(n / (4 * lanes)))
  uint8_t lanes[PKA_MAX_REDUCE_JOBS];              // float4 lanes per CTA of the job: 32, 64 or 128
  int32_t n_jobs;
};

// A CTA owns `lanes * 4` consecutive output elements of one job (lanes = 32, 64 or 128, chosen per job by the host:
// few lanes when there are many partials to sum, so that the 256 threads split the *partials* instead of idling):
// thread (g, q) = (threadIdx.x / lanes, threadIdx.x % lanes) sums the partials s = g, g + G, g + 2G, ... (G = 256 / lanes)
// of float4 q with eight independent loads in flight; the G group sums are then added through shared memory in the
// fixed order g = 0 .. G-1 (bit-reproducible).  Partial s lives at src + s * split_stride.
__global__ void __launch_bounds__(256)
reduce_jobs_kernel(const __grid_constant__ ReduceTable tab) {
  pdl_wait();
  __shared__ float4 red[256];
  // locate the job of this CTA (n_jobs <= 128: binary search of the prefix table, CTA-uniform)
  int j = 0;
  {
    int lo = 0, hi = tab.n_jobs;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tab.cta_start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid; }
    j = lo;
  }
  const pka_reduce_job& jb = tab.job[j];
  const int lanes = tab.lanes[j], G = 256 / lanes;
  const int g = threadIdx.x / lanes, q = threadIdx.x % lanes;
  const long long e0 = ((long long)(blockIdx.x - tab.cta_start[j]) * (lanes * 4)) + q * 4;
  float4 acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool in = e0 < jb.n;
  const bool vec = in && e0 + 4 <= jb.n && ((jb.split_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(jb.src) & 15u) == 0);
  if (in) {
    const float* base = jb.src + e0;
    int s = g;
    if (vec) {
      for (; s + 7 * G < jb.splits; s += 8 * G) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float4*>(base + (long long)(s + u * G) * jb.split_stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc[u & 3].x += v[u].x; acc[u & 3].y += v[u].y; acc[u & 3].z += v[u].z; acc[u & 3].w += v[u].w; }
      }
      for (; s < jb.splits; s += G) {
        const float4 v = *reinterpret_cast<const float4*>(base + (long long)s * jb.split_stride);
        acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
      }
    } else {
      const int left = (int)min((long long)4, jb.n - e0);
      for (; s < jb.splits; s += G) {
        const float* pp = base + (long long)s * jb.split_stride;
        acc[0].x += pp[0];
        if (left > 1) acc[0].y += pp[1];
        if (left > 2) acc[0].z += pp[2];
        if (left > 3) acc[0].w += pp[3];
      }
    }
  }
  float4 t = make_float4((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y),
                         (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z), (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w));
  red[threadIdx.x] = t;
  __syncthreads();
  if (g == 0 && in) {
    for (int k = 1; k < G; ++k) { const float4 o = red[k * lanes + q]; t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w; }
    const float tv[4] = {t.x, t.y, t.z, t.w};
    const int left = (int)min((long long)4, jb.n - e0);
    if (jb.kind == PKA_REDUCE_PLAIN) {
      float* dst = jb.dst + e0;
      if (left == 4 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        if (jb.accumulate) { const float4 c = *reinterpret_cast<float4*>(dst); t.x += c.x; t.y += c.y; t.z += c.z; t.w += c.w; }
        *reinterpret_cast<float4*>(dst) = t;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) if (i < left) dst[i] = jb.accumulate ? dst[i] + tv[i] : tv[i];
      }
    } else {
      // PKA_REDUCE_HEADS: src index e = (h*dk + j)*D + d  (one packed head block [(h,j), d])  ->  dst[(h*D + d)*dk + j]
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i >= left) break;
        const long long e = e0 + i;
        const int d = (int)(e % jb.D);
        const int n = (int)(e / jb.D);
        const int jj = n % jb.dk, h = n / jb.dk;
        jb.dst[((long long)h * jb.D + d) * jb.dk + jj] = tv[i];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- operand relayout jobs
struct RelayoutTable {
  pka_relayout_job job[PKA_MAX_RELAYOUT_JOBS];
  int32_t cta_start[PKA_MAX_RELAYOUT_JOBS + 1];
  int32_t n_jobs;
  unsigned long long* counter;                     // optional: the model's dropout step counter, advanced by one
};
constexpr int kRelElemsPerCta = 2048;

// kind PLAIN: W fp32 [N, nseg*K] -> Wf bf16 [N, ldf] (same element order, row pitch ldf) and/or
//             Wd bf16 [K, ldd] with Wd[i, s*N + o] = W[o, s*K + i]  (data-gradient operand)
// kind HEADS: w fp32 [H, D, dk] (T/SubLayers.py:29-31) -> rows n0 + h*dk + j of Wf [.., ldf] (Wf[n, d]) and columns
//             n0 + h*dk + j of Wd [D, ldd] (Wd[d, n]): one block of a packed q|k|v (or k|v of several layers) operand
__global__ void __launch_bounds__(256)
relayout_jobs_kernel(const __grid_constant__ RelayoutTable tab) {
  pdl_wait();
  if (tab.counter && blockIdx.x == 0 && threadIdx.x == 0) tab.counter[0] += 1ull;
  if (tab.n_jobs == 0) return;
  int j = 0;
  {
    int lo = 0, hi = tab.n_jobs;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tab.cta_start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid; }
    j = lo;
  }
  const pka_relayout_job& jb = tab.job[j];
  const long long total = jb.kind == PKA_RELAYOUT_PLAIN ? (long long)jb.N * jb.nseg * jb.K : (long long)jb.N * jb.K * jb.nseg;
  const long long e_lo = (long long)(blockIdx.x - tab.cta_start[j]) * kRelElemsPerCta;
  __nv_bfloat16* Wf = (__nv_bfloat16*)jb.wf;
  __nv_bfloat16* Wd = (__nv_bfloat16*)jb.wd;
#pragma unroll 2
  for (int it = 0; it < kRelElemsPerCta / 256; ++it) {
    const long long e = e_lo + it * 256 + threadIdx.x;
    if (e >= total) break;
    const __nv_bfloat16 v = __float2bfloat16_rn(jb.src[e]);
    if (jb.kind == PKA_RELAYOUT_PLAIN) {
      const int row_len = jb.nseg * jb.K;
      const int o = (int)(e / row_len), rem = (int)(e % row_len);
      if (Wf) Wf[(long long)o * jb.ldf + rem] = v;
      if (Wd) { const int s = rem / jb.K, i = rem % jb.K; Wd[(long long)i * jb.ldd + (long long)s * jb.N + o] = v; }
    } else {                                       // N = H, K = D, nseg = dk
      const int dk = jb.nseg, D = jb.K;
      const int h = (int)(e / ((long long)D * dk)), d = (int)((e / dk) % D), jj = (int)(e % dk);
      const int n = jb.n0 + h * dk + jj;
      if (Wf) Wf[(long long)n * jb.ldf + d] = v;
      if (Wd) Wd[(long long)d * jb.ldd + n] = v;
    }
  }
}

// teacher-forcing split (L/train.py:163-165): tgt[B, L1] -> tgt_in = tgt[:, :-1], goal = tgt[:, 1:], mask_in = mask[:, :-1]
__global__ void split_targets_kernel(const long long* __restrict__ tgt, const uint8_t* __restrict__ mask,
                                     long long* __restrict__ tgt_in, long long* __restrict__ goal,
                                     uint8_t* __restrict__ mask_in, int B, int L1) {
  pdl_wait();
  const int L = L1 - 1;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < B * L; e += gridDim.x * blockDim.x) {
    const int b = e / L, i = e % L;
    const long long* row = tgt + (long long)b * L1;
    tgt_in[e] = row[i];
    goal[e] = row[i + 1];
    mask_in[e] = mask[(long long)b * L1 + i];
  }
}

// [rows, N] fp32 or bf16 -> bf16 [rows, Np] with the columns N..Np-1 zero: the gradient of an output whose width is not
// a multiple of 8 (the vocabulary projection, V = 53) as a TMA-legal operand (16-byte row pitch) of the tensor-core GEMMs
template <typename T>
__global__ void pad_cast_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ y, long long rows, int N, int Np) {
  pdl_wait();
  const long long total = rows * Np;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / Np;
    const int c = (int)(e - r * Np);
    y[e] = c < N ? __float2bfloat16_rn(to_f(x[r * N + c])) : __float2bfloat16_rn(0.f);
  }
}

}  // namespace pka

using namespace pka;

extern "C" int pka_pad_cast(const void* x, int dtype, void* y, int64_t rows, int N, int Np, void* stream) {
  PKA_REQUIRE(x && y && rows > 0 && N > 0 && Np >= N, PKA_EINVAL, "pad_cast: bad arguments");
  const long long total = rows * Np;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
  if (dtype == PKA_F32) launch_k(pad_cast_kernel<float>, (int)blocks, 256, 0, as_stream(stream), (const float*)x, (__nv_bfloat16*)y, (long long)rows, N, Np);
  else if (dtype == PKA_BF16) launch_k(pad_cast_kernel<__nv_bfloat16>, (int)blocks, 256, 0, as_stream(stream), (const __nv_bfloat16*)x, (__nv_bfloat16*)y, (long long)rows, N, Np);
  else PKA_REQUIRE(false, PKA_EUNSUPPORTED, "pad_cast: dtype %d", dtype);
  return check_launch("pad_cast");
}

extern "C" int pka_split_targets(const int64_t* tgt, const uint8_t* mask, int64_t* tgt_in, int64_t* goal, uint8_t* mask_in,
                                 int B, int L1, void* stream) {
  PKA_REQUIRE(tgt && mask && tgt_in && goal && mask_in && B > 0 && L1 > 1, PKA_EINVAL, "split_targets: bad arguments");
  const int total = B * (L1 - 1);
  launch_k(split_targets_kernel, (total + 255) / 256, 256, 0, as_stream(stream), (const long long*)tgt, mask, (long long*)tgt_in,
           (long long*)goal, mask_in, B, L1);
  return check_launch("split_targets");
}

extern "C" int pka_reduce_jobs(const pka_reduce_job* jobs, int n_jobs, void* stream) {
  PKA_REQUIRE(jobs && n_jobs >= 0, PKA_EINVAL, "reduce_jobs: null table");
  for (int base = 0; base < n_jobs; base += PKA_MAX_REDUCE_JOBS) {
    ReduceTable tab;
    const int n = n_jobs - base < PKA_MAX_REDUCE_JOBS ? n_jobs - base : PKA_MAX_REDUCE_JOBS;
    int ctas = 0;
    for (int i = 0; i < n; ++i) {
      const pka_reduce_job& jb = jobs[base + i];
      PKA_REQUIRE(jb.src && jb.dst && jb.n > 0 && jb.splits >= 1, PKA_EINVAL, "reduce_jobs: job %d: bad arguments", base + i);
      PKA_REQUIRE(jb.kind == PKA_REDUCE_PLAIN || (jb.kind == PKA_REDUCE_HEADS && jb.D > 0 && jb.dk > 0 && jb.n % ((int64_t)jb.D * jb.dk) == 0),
                  PKA_EINVAL, "reduce_jobs: job %d: bad kind / head geometry", base + i);
      tab.job[i] = jb;
      tab.cta_start[i] = ctas;
      const int lanes = jb.splits >= 64 ? 32 : (jb.splits >= 16 ? 64 : 128);
      tab.lanes[i] = (uint8_t)lanes;
      ctas += (int)((jb.n + lanes * 4 - 1) / (lanes * 4));
    }
    tab.cta_start[n] = ctas;
    tab.n_jobs = n;
    if (ctas == 0) continue;
    launch_k(reduce_jobs_kernel, ctas, 256, 0, as_stream(stream), tab);
    int rc = check_launch("reduce_jobs");
    if (rc) return rc;
  }
  return PKA_OK;
}

extern "C" int pka_relayout_jobs(const pka_relayout_job* jobs, int n_jobs, uint64_t* step_counter, void* stream) {
  PKA_REQUIRE((jobs || n_jobs == 0) && n_jobs >= 0, PKA_EINVAL, "relayout_jobs: null table");
  bool ticked = false;
  for (int base = 0; base < n_jobs || (!ticked && step_counter); base += PKA_MAX_RELAYOUT_JOBS) {
    RelayoutTable tab;
    const int left = n_jobs - base;
    const int n = left < PKA_MAX_RELAYOUT_JOBS ? (left > 0 ? left : 0) : PKA_MAX_RELAYOUT_JOBS;
    int ctas = 0;
    for (int i = 0; i < n; ++i) {
      const pka_relayout_job& jb = jobs[base + i];
      PKA_REQUIRE(jb.src && (jb.wf || jb.wd) && jb.N > 0 && jb.K > 0 && jb.nseg >= 1, PKA_EINVAL, "relayout_jobs: job %d: bad arguments", base + i);
      PKA_REQUIRE(jb.kind == PKA_RELAYOUT_PLAIN || jb.kind == PKA_RELAYOUT_HEADS, PKA_EINVAL, "relayout_jobs: job %d: bad kind", base + i);
      tab.job[i] = jb;
      tab.cta_start[i] = ctas;
      const long long total = (long long)jb.N * jb.K * jb.nseg;
      ctas += (int)((total + kRelElemsPerCta - 1) / kRelElemsPerCta);
    }
    tab.cta_start[n] = ctas;
    tab.n_jobs = n;
    tab.counter = ticked ? nullptr : (unsigned long long*)step_counter;
    ticked = true;
    if (ctas == 0) { if (!tab.counter) continue; ctas = 1; tab.n_jobs = 0; }
    launch_k(relayout_jobs_kernel, ctas, 256, 0, as_stream(stream), tab);
    int rc = check_launch("relayout_jobs");
    if (rc) return rc;
  }
  return PKA_OK;
}
