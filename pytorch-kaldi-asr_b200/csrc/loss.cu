// Summed cross-entropy with PAD mask, optional label smoothing, fused arg-max accuracy (forward) and
// (softmax - target) gradient (backward).  Replaces cal_loss / get_performance, L/train.py:58-90:
//   plain:    loss = sum_{goal!=PAD} -log_softmax(x)[goal]                       (F.cross_entropy(ignore_index=0, 'sum'))
//   smoothed: target = 1-eps on goal, eps/(V-1) on every other class (PAD/UNK/BOS columns included), eps = 0.1
//   n_correct = #{goal!=PAD and argmax(x)==goal}  (first maximal index, like torch.max on CPU)
// One warp per row; the final sum over rows is a fixed-order two-stage reduction (deterministic, no atomics).
#include "common.cuh"
#include <math_constants.h>

namespace pka {

constexpr int kCeWarps = 8;

template <typename T>
__global__ void __launch_bounds__(kCeWarps * 32)
ce_fwd_kernel(const T* __restrict__ logits, const long long* __restrict__ goal, int N, int V, int smoothing,
              float eps, float* __restrict__ lse_o, float* part, unsigned int* done, float* out3) {
  pdl_wait();
  __shared__ float red[kCeWarps][3];
  __shared__ bool s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float loss = 0.f, correct = 0.f, words = 0.f;
  for (int row = blockIdx.x * kCeWarps + warp; row < N; row += gridDim.x * kCeWarps) {
    const T* xr = logits + (long long)row * V;
    float mx = -CUDART_INF_F;
    int arg = 0x7fffffff;
    for (int c = lane; c < V; c += 32) {
      const float v = to_f(xr[c]);
      if (v > mx) { mx = v; arg = c; }                 // strict > keeps the first index within a lane
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    float se = 0.f, sx = 0.f;
    for (int c = lane; c < V; c += 32) {
      const float v = to_f(xr[c]);
      se += expf(v - mx);
      sx += v;
    }
    se = warp_sum(se); sx = warp_sum(sx);
    const float lse = mx + logf(se);
    const long long g = goal[row];
    if (lane == 0) {
      lse_o[row] = lse;
      if (g != 0) {
        const float picked = to_f(xr[g]) - lse;          // log p[goal]
        if (smoothing) {
          const float off = eps / (float)(V - 1);
          const float all = sx - (float)V * lse;         // sum_c log p[c]
          loss += -((1.f - eps) * picked + off * (all - picked));
        } else {
          loss += -picked;
        }
        words += 1.f;
        if ((long long)arg == g) correct += 1.f;
      }
    }
  }
  if (lane == 0) { red[warp][0] = loss; red[warp][1] = correct; red[warp][2] = words; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int w = 0; w < kCeWarps; ++w) s += red[w][threadIdx.x];
    part[blockIdx.x * 3 + threadIdx.x] = s;
  }
  if (!done) return;
  // fused finish: the LAST CTA to arrive sums the per-CTA partials in index order (the same fixed order as
  // ce_finish_kernel, whichever CTA happens to be last), so the separate single-CTA launch disappears
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int k = threadIdx.x >> 5, ln = threadIdx.x & 31;
  if (k < 3) {
    float s = 0.f;
    for (int b = ln; b < (int)gridDim.x; b += 32) s += __ldcg(part + b * 3 + k);     // L2: written by other SMs
    s = warp_sum(s);
    if (ln == 0) out3[k] = s;
  }
  if (threadIdx.x == 0) *done = 0u;
}

__global__ void ce_finish_kernel(const float* __restrict__ part, int nblk, float* __restrict__ out3) {
  pdl_wait();
  // three warps, one per statistic: lane l sums the CTAs l, l+32, ... in index order, then a shuffle tree
  // (a fixed order => run-to-run identical; round 1 used three serial lanes: ~7 us for 252 partial rows)
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (k < 3) {
    float s = 0.f;
    for (int b = lane; b < nblk; b += 32) s += part[b * 3 + k];
    s = warp_sum(s);
    if (lane == 0) out3[k] = s;
  }
}

template <typename T>
__global__ void ce_bwd_kernel(const T* __restrict__ logits, const long long* __restrict__ goal,
                              const float* __restrict__ lse, const float* __restrict__ grad_out, T* __restrict__ dl,
                              int N, int V, int smoothing, float eps) {
  pdl_wait();
  const float go = grad_out ? grad_out[0] : 1.f;
  const float off = smoothing ? eps / (float)(V - 1) : 0.f;
  const float on = smoothing ? 1.f - eps : 1.f;
  const long long total = (long long)N * V;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int row = e / V, c = e % V;
    const long long g = goal[row];
    float v = 0.f;
    if (g != 0) {
      const float pr = expf(to_f(logits[e]) - lse[row]);
      // d/dx of -sum_c t_c log p_c = p*sum(t) - t ; sum(t) = 1 for both target distributions
      v = go * (pr - (c == g ? on : off));
    }
    dl[e] = from_f<T>(v);
  }
}

}  // namespace pka

extern "C" int pka_ce_blocks(int N) {
  int need = (N + pka::kCeWarps - 1) / pka::kCeWarps;
  int cap = pka::kNumSMs * 2;
  return need < cap ? (need > 0 ? need : 1) : cap;
}

extern "C" int pka_ce_fwd(const void* logits, const int64_t* goal, int dtype, int N, int V, int smoothing, float eps,
                          float* out3, float* lse, float* part_ws, void* stream) {
  using namespace pka;
  PKA_REQUIRE(logits && goal && out3 && lse && part_ws, PKA_EINVAL, "ce_fwd: null pointer");
  PKA_REQUIRE(N > 0 && V > 1, PKA_EINVAL, "ce_fwd: N=%d V=%d", N, V);
  const int nblk = pka_ce_blocks(N);
  cudaStream_t st = as_stream(stream);
  if (dtype == PKA_F32)
    launch_k(ce_fwd_kernel<float>, nblk, kCeWarps * 32, 0, st, (const float*)logits, (const long long*)goal, N, V, smoothing, eps, lse, part_ws, (unsigned int*)nullptr, (float*)nullptr);
  else if (dtype == PKA_BF16)
    launch_k(ce_fwd_kernel<__nv_bfloat16>, nblk, kCeWarps * 32, 0, st, (const __nv_bfloat16*)logits, (const long long*)goal, N, V, smoothing, eps, lse, part_ws, (unsigned int*)nullptr, (float*)nullptr);
  else PKA_REQUIRE(false, PKA_EUNSUPPORTED, "ce_fwd: dtype %d", dtype);
  int rc = check_launch("ce_fwd");
  if (rc) return rc;
  launch_k(ce_finish_kernel, 1, 96, 0, st, part_ws, nblk, out3);
  return check_launch("ce_finish");
}

// pka_ce_fwd with the fixed-order finish done by the last CTA of the same launch.  done_counter: uint32[1], zero before
// the first call (resets itself); one counter per stream that may run the loss concurrently.
extern "C" int pka_ce_fwd_fused(const void* logits, const int64_t* goal, int dtype, int N, int V, int smoothing, float eps,
                                float* out3, float* lse, float* part_ws, uint32_t* done_counter, void* stream) {
  using namespace pka;
  PKA_REQUIRE(logits && goal && out3 && lse && part_ws && done_counter, PKA_EINVAL, "ce_fwd_fused: null pointer");
  PKA_REQUIRE(N > 0 && V > 1, PKA_EINVAL, "ce_fwd_fused: N=%d V=%d", N, V);
  const int nblk = pka_ce_blocks(N);
  cudaStream_t st = as_stream(stream);
  if (dtype == PKA_F32)
    launch_k(ce_fwd_kernel<float>, nblk, kCeWarps * 32, 0, st, (const float*)logits, (const long long*)goal, N, V, smoothing, eps, lse, part_ws, (unsigned int*)done_counter, out3);
  else if (dtype == PKA_BF16)
    launch_k(ce_fwd_kernel<__nv_bfloat16>, nblk, kCeWarps * 32, 0, st, (const __nv_bfloat16*)logits, (const long long*)goal, N, V, smoothing, eps, lse, part_ws, (unsigned int*)done_counter, out3);
  else PKA_REQUIRE(false, PKA_EUNSUPPORTED, "ce_fwd_fused: dtype %d", dtype);
  return check_launch("ce_fwd_fused");
}

extern "C" int pka_ce_bwd(const void* logits, const int64_t* goal, const float* lse, const float* grad_out,
                          void* dlogits, int dtype, int N, int V, int smoothing, float eps, void* stream) {
  using namespace pka;
  PKA_REQUIRE(logits && goal && lse && dlogits, PKA_EINVAL, "ce_bwd: null pointer");
  PKA_REQUIRE(N > 0 && V > 1, PKA_EINVAL, "ce_bwd: N=%d V=%d", N, V);
  const long long total = (long long)N * V;
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  cudaStream_t st = as_stream(stream);
  if (dtype == PKA_F32)
    launch_k(ce_bwd_kernel<float>, blocks, 256, 0, st, (const float*)logits, (const long long*)goal, lse, grad_out, (float*)dlogits, N, V, smoothing, eps);
  else if (dtype == PKA_BF16)
    launch_k(ce_bwd_kernel<__nv_bfloat16>, blocks, 256, 0, st, (const __nv_bfloat16*)logits, (const long long*)goal, lse, grad_out, (__nv_bfloat16*)dlogits, N, V, smoothing, eps);
  else PKA_REQUIRE(false, PKA_EUNSUPPORTED, "ce_bwd: dtype %d", dtype);
  return check_launch("ce_bwd");
}
