// Fused front-end: [per-utterance CMVN] -> frame folding (sub-sampling by stacking) -> frame splicing, one pass over
// the padded feature batch, 128-bit coalesced accesses.  HBM-bound:
//   bytes = B*T*F*4 (read, the spliced re-reads hit L1/L2) + B*(T/fold)*(n_ctx*F*fold)*sizeof(out) (write).
// Replaces: Kaldi apply-cmvn (P/run.sh:37-42, external binary) + fold_seq_and_mask (T/Models.py:51-65) +
// ConcatLayer (L/pytorch/TDNN.py:20-28).  Splicing pads with zeros beyond the *tensor* edge (row index outside
// [0, T/fold)), exactly like ConcatLayer's F.pad; padded frames inside the tensor are whatever the input holds (zeros),
// and stay zero under CMVN because the reference pads after feature extraction.
#include "common.cuh"

namespace pka {

// stats[b][0][f] = mean, stats[b][1][f] = 1/sqrt(var) (or 1 when only the mean is removed)
__global__ void __launch_bounds__(256)
cmvn_stats_kernel(const float* __restrict__ x, const int* __restrict__ lengths, float* __restrict__ stats, int T, int F,
                  int norm_vars) {
  extern __shared__ double sm[];                  // [8 warps][2][F]
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int n = lengths[b];
  n = n < 0 ? 0 : (n > T ? T : n);
  const float* xb = x + (long long)b * T * F;
  for (int f0 = 0; f0 < F; f0 += 32) {
    const int f = f0 + lane;
    double s = 0.0, s2 = 0.0;
    if (f < F)
      for (int t = warp; t < n; t += 8) { const double v = xb[(long long)t * F + f]; s += v; s2 += v * v; }
    if (f < F) { sm[(warp * 2 + 0) * F + f] = s; sm[(warp * 2 + 1) * F + f] = s2; }
  }
  __syncthreads();
  for (int f = threadIdx.x; f < F; f += 256) {
    double s = 0.0, s2 = 0.0;
    for (int w = 0; w < 8; ++w) { s += sm[(w * 2 + 0) * F + f]; s2 += sm[(w * 2 + 1) * F + f]; }
    double mean = n > 0 ? s / n : 0.0, istd = 1.0;
    if (norm_vars && n > 0) {
      double var = s2 / n - mean * mean;
      if (var < 1e-20) var = 1e-20;
      istd = 1.0 / sqrt(var);
    }
    stats[((long long)b * 2 + 0) * F + f] = (float)mean;
    stats[((long long)b * 2 + 1) * F + f] = (float)istd;
  }
}

struct FrontP {
  int B, T, F, fold, n_ctx, cmvn;
  int ctx[PKA_MAX_CTX];
};

template <typename To, int VEC>
__global__ void __launch_bounds__(256)
frontend_kernel(const FrontP p, const float* __restrict__ x, const int* __restrict__ lengths,
                const float* __restrict__ stats, To* __restrict__ out) {
  const int Tf = p.T / p.fold, Ff = p.F * p.fold, W = p.n_ctx * Ff;
  const long long total = (long long)p.B * Tf * (W / VEC);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(e % (W / VEC)) * VEC;
    const long long row = e / (W / VEC);
    const int t = (int)(row % Tf), b = (int)(row / Tf);
    const int c = col / Ff, fp = col % Ff;
    const int ts = t + p.ctx[c];
    float v[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = 0.f;
    if (ts >= 0 && ts < Tf) {
      const int frame = ts * p.fold + fp / p.F, f = fp % p.F;       // VEC consecutive f stay inside one frame (F%VEC==0)
      const float* src = x + ((long long)b * p.T + frame) * p.F + f;
      if (VEC == 4) { const float4 q = *reinterpret_cast<const float4*>(src); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
      else v[0] = src[0];
      if (p.cmvn) {
        const bool real = frame < lengths[b];
        const float* mu = stats + ((long long)b * 2 + 0) * p.F + f;
        const float* is = stats + ((long long)b * 2 + 1) * p.F + f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = real ? (v[i] - mu[i]) * is[i] : 0.f;
      }
    }
    To* dst = out + row * W + col;
    if (VEC == 4) st4(dst, make_float4(v[0], v[1], v[2], v[3]));
    else dst[0] = from_f<To>(v[0]);
  }
}

}  // namespace pka

extern "C" int pka_frontend_fwd(const float* feats, const int32_t* lengths, void* out, int out_dtype, int B, int T, int F,
                                int fold, const int32_t* ctx_host, int n_ctx, int cmvn_mode, float* stats_ws,
                                void* stream) {
  using namespace pka;
  PKA_REQUIRE(feats && out && ctx_host, PKA_EINVAL, "frontend_fwd: null pointer");
  PKA_REQUIRE(B > 0 && T > 0 && F > 0 && fold >= 1 && T / fold >= 1, PKA_EINVAL, "frontend_fwd: B=%d T=%d F=%d fold=%d", B, T, F, fold);
  PKA_REQUIRE(n_ctx >= 1 && n_ctx <= PKA_MAX_CTX, PKA_EUNSUPPORTED, "frontend_fwd: n_ctx=%d (max %d)", n_ctx, PKA_MAX_CTX);
  PKA_REQUIRE(cmvn_mode >= 0 && cmvn_mode <= 2, PKA_EINVAL, "frontend_fwd: cmvn_mode=%d", cmvn_mode);
  PKA_REQUIRE(cmvn_mode == 0 || (lengths && stats_ws), PKA_EINVAL, "frontend_fwd: CMVN needs lengths and stats_ws");
  cudaStream_t st = as_stream(stream);
  if (cmvn_mode) {
    cmvn_stats_kernel<<<B, 256, 8 * 2 * F * sizeof(double), st>>>(feats, lengths, stats_ws, T, F, cmvn_mode == 2);
    int rc = check_launch("cmvn_stats");
    if (rc) return rc;
  }
  FrontP p;
  p.B = B; p.T = T; p.F = F; p.fold = fold; p.n_ctx = n_ctx; p.cmvn = cmvn_mode;
  for (int i = 0; i < PKA_MAX_CTX; ++i) p.ctx[i] = i < n_ctx ? ctx_host[i] : 0;
  const bool vec = (F % 4 == 0) && aligned16(feats) && aligned16(out);
  const long long W = (long long)n_ctx * F * fold;
  const long long total = (long long)B * (T / fold) * (vec ? W / 4 : W);
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
  if (out_dtype == PKA_F32) {
    if (vec) frontend_kernel<float, 4><<<(int)blocks, 256, 0, st>>>(p, feats, lengths, stats_ws, (float*)out);
    else frontend_kernel<float, 1><<<(int)blocks, 256, 0, st>>>(p, feats, lengths, stats_ws, (float*)out);
  } else if (out_dtype == PKA_BF16) {
    if (vec) frontend_kernel<__nv_bfloat16, 4><<<(int)blocks, 256, 0, st>>>(p, feats, lengths, stats_ws, (__nv_bfloat16*)out);
    else frontend_kernel<__nv_bfloat16, 1><<<(int)blocks, 256, 0, st>>>(p, feats, lengths, stats_ws, (__nv_bfloat16*)out);
  } else PKA_REQUIRE(false, PKA_EUNSUPPORTED, "frontend_fwd: out dtype %d", out_dtype);
  return check_launch("frontend");
}
