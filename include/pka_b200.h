/*
 * pka_b200.h -- C ABI of libpka_b200.so, the sm_100a kernel library under the Python mirror of
 * boji123/pytorch-kaldi-asr's acoustic-model hot path.
 *
 * The reference has no FFI of its own: its "kernel layer" is stock ATen called from Python modules.  Each entry
 * point below therefore names the reference *op sequence* (file:line under /root/reference) it replaces.
 *   T/ = project/attention-transformer-timit/local/pytorch/transformer/
 *   L/ = project/attention-transformer-timit/local/      U/ = pytorch/utils/
 *
 * Contract (all entry points)
 *   - plain C: raw device pointers, sizes, a cudaStream_t passed as void*; no torch types.
 *   - the caller (PyTorch) owns every buffer; the library never allocates, frees or keeps a pointer.
 *   - all work is enqueued on `stream`; no host synchronisation; CUDA-graph capturable.
 *   - return 0 on success, a PKA_E* code otherwise; pka_last_error() gives a thread-local message.
 *   - `dtype`: element type of activations in HBM (PKA_F32 / PKA_BF16); parameters and statistics are fp32.
 *   - dropout: keep-mask bit of element `i` at `site` in training step `*step_ptr` is a pure function
 *     philox4x32-10(seed, site, step, i) so the backward pass regenerates it instead of storing it;
 *     p == 0 disables it.  pka_dropout_mask() materialises the same bits for tests.
 */
#ifndef PKA_B200_H
#define PKA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { PKA_F32 = 0, PKA_BF16 = 1 };
enum {
  PKA_OK = 0,
  PKA_EINVAL = 1,      /* bad shape / argument                      */
  PKA_EUNSUPPORTED = 2,/* dimension or dtype outside what is built  */
  PKA_EALIGN = 3,      /* pointer or leading dimension misaligned   */
  PKA_ELAUNCH = 4,     /* cudaGetLastError() after launch           */
  PKA_EDEVICE = 5      /* not an sm_100 device / driver entry point missing */
};

#define PKA_MAX_CTX 8

typedef struct {
  float p;                         /* drop probability; 0 = off */
  uint32_t site;                   /* dropout site id (one per nn.Dropout call site of the reference) */
  uint64_t seed;
  const uint64_t* step_ptr;        /* device counter, advanced once per training step (may be NULL = 0) */
} pka_dropout;

const char* pka_last_error(void);
int pka_version(void);
/* number of kernels this library has launched so far in this process (one per successful launch call) */
uint64_t pka_launch_count(void);
/* 0 if the current device is compute capability 10.x, PKA_EDEVICE otherwise */
int pka_check_device(void);

/* ---- (a) front-end: [optional per-utterance CMVN] -> frame folding -> frame splicing --------------------------
 * replaces: Kaldi `apply-cmvn` (P/run.sh:37-42) + fold_seq_and_mask (T/Models.py:51-65) + ConcatLayer
 * (L/pytorch/TDNN.py:20-28).  feats f32[B,T,F] zero padded; lengths[B] (only read when cmvn_mode != 0);
 * out[B, T/fold, n_ctx*F*fold]; zero rows are spliced in beyond the *tensor* edges exactly like ConcatLayer.
 * cmvn_mode: 0 none, 1 mean, 2 mean+variance.  stats_ws: float[B*2*F] scratch (cmvn only). */
int pka_frontend_fwd(const float* feats, const int32_t* lengths, void* out, int out_dtype, int B, int T, int F,
                     int fold, const int32_t* ctx_host, int n_ctx, int cmvn_mode, float* stats_ws, void* stream);

/* ---- (c) GEMM family, fp32 SIMT "exact" path -------------------------------------------------------------------
 * C[M,N] (+)= epilogue( sum_{s<nseg} sum_{k<K} A_s(m,k) * B_s(k,n) ), batched over `nbatch` (grid.z).
 *   transA=0: A_s(m,k) = A[(m+shiftA[s])*lda + s*a_seg_off + k]; rows are (utterance, frame) pairs with T frames per
 *             utterance, a shifted row that leaves [0,T) reads as 0  -> implicit ConcatLayer, never materialised.
 *   transA=1: A(m,k)   = A[k*lda + m]                               (weight-gradient: reduction over frames)
 *   transB=1: B_s(k,n) = B[n*ldb + s*b_seg_off + k]                 (nn.Linear weight [out,in])
 *   transB=0: B_s(k,n) = B[(k+shiftB[batch])*ldb + s*b_seg_off + n] (frame shift on the reduction index, same T rule)
 * epilogue: + bias[n], ReLU, dropout, + residual[m*ldr+n], accumulate into C.
 * replaces: BottleLinear (T/Modules.py:8-30), TDNNLayer cat+Linear+ReLU+dropout (L/pytorch/TDNN.py:41-46),
 * LDALayer (:53-55), per-head bmm projections (T/SubLayers.py:49-56), Conv1d k=1 FFN (T/SubLayers.py:81-83), and
 * their autograd backward. */
typedef struct {
  const void* A; const void* B; void* C;
  const float* bias; const void* residual;
  int32_t M, N, K, nseg, nbatch;
  int32_t lda, ldb, ldc, ldr;
  int32_t transA, transB;
  int64_t a_seg_off, b_seg_off;
  int64_t a_batch_off, b_batch_off, c_batch_off;
  int32_t shiftA[PKA_MAX_CTX];
  int32_t shiftB[PKA_MAX_CTX];
  int32_t T;
  int32_t relu;
  int32_t accumulate;
  int32_t splitk;          /* >1: split the reduction over `splitk` CTAs per tile (nseg==1, no epilogue); partial */
  pka_dropout drop;        /*     sums go to splitk_ws and are added in a fixed order (deterministic)             */
  void* splitk_ws;         /* float[splitk*nbatch*M*N] */
} pka_gemm_desc;
int pka_gemm_f32(const pka_gemm_desc* d, void* stream);

/* ---- (c) GEMM family, bf16 tcgen05 / TMEM / TMA path ------------------------------------------------------------
 * UMMA 128x128x16 tiles (tcgen05.mma, fp32 accumulation in TMEM), operands staged by TMA with 128B swizzle through a
 * 4-stage mbarrier ring.  Activations are described to TMA as 3-D tensors [utterance, frame, feature], so the frame
 * splice of a TDNN layer is `nseg` shifted boxes accumulating into one TMEM tile and TMA's out-of-bounds zero fill IS
 * ConcatLayer's zero padding (L/pytorch/TDNN.py:20-28).
 *   mode 0: C[b,t,:] = epi( sum_seg A[b, t+shift[seg], seg*a_seg_col : +K] . B[:, seg*b_seg_col : +K]^T )
 *           A bf16 [Bt,T,lda], B bf16 [N,ldb] (nn.Linear layout or pka_weight_relayout's data-gradient layout),
 *           C bf16/fp32 [Bt*T, ldc]; epilogue + bias[n], ReLU, dropout; Ct (optional) = bf16 transposed copy
 *           [N, Bt, Tp] (Tp = T rounded up to 8) that mode 1 consumes.
 *   mode 1: weight gradient  dW[o, seg*N + i] = sum_{b,t} A[o,b,t] * B[i,b,t+shift[seg]]   with A = dZt [M,Bt,Tp],
 *           B = Xt [N,Bt,Tp] (both transposed activations, bf16); the frame reduction is split over `splits` CTAs
 *           per tile, C = fp32 partials [splits, M, ldc] (ldc = nseg*N), summed by pka_tc_reduce in fixed order.
 * replaces: TDNNLayer / BottleLinear / LDALayer GEMMs and their backward in the bf16 training path. */
typedef struct {
  const void* A; const void* B; void* C; void* Ct;
  const float* bias;
  int32_t mode, Bt, T, Tp;
  int32_t M, N, K, nseg;
  int32_t lda, ldb, ldc;
  int32_t a_seg_col, b_seg_col;
  int32_t shift[PKA_MAX_CTX];
  int32_t relu, c_dtype, splits, reserved;
  pka_dropout drop;
  const void* addend;      /* mode 0, optional: bf16 [Bt*T, ldadd], C = epi(...) + addend -- the residual branch of a      */
  int32_t ldadd, reserved2;/* gradient (dX = dZ.W + dResidual) folded into the data-gradient GEMM's epilogue             */
} pka_tc_desc;
int pka_gemm_tc(const pka_tc_desc* d, void* stream);
/* n_desc independent mode-2 (weight-gradient) problems in ONE launch: the decoder-shaped gradients of a backward pass
 * (a few output tiles each, ~6 us when launched alone) are not on the critical path, so they are collected and run as
 * one grid at the end of backward.  Same partial layout as pka_gemm_tc mode 2 (C = fp32 [splits, M, ldc]). */
int pka_gemm_tc_wgrad_group(const pka_tc_desc* descs_host, int n_desc, void* stream);
/* out[e] (+)= sum_{s<splits} ws[s*per + e] */
int pka_tc_reduce(const float* ws, float* out, int64_t per, int splits, int accumulate, void* stream);
/* fp32 W[N, nseg*K] -> Wf bf16 (same layout, forward operand) and/or Wd bf16 [K, nseg*N], Wd[i, s*N+o] = W[o, s*K+i] */
int pka_weight_relayout(const float* W, void* Wf, void* Wd, int N, int K, int nseg, void* stream);
/* per-head projection weights w_p[H,D,dk], p < P <= 3 (T/SubLayers.py:29-31) -> packed bf16 GEMM operands
 * Wf[(p*H+h)*dk+j, d] (forward) and Wd[d, (p*H+h)*dk+j] (data-gradient); and the way back for the weight gradient
 * dWcat fp32 [(p,h,j), d] -> g_p[h,d,j]. */
int pka_head_weight_relayout(const float* w0, const float* w1, const float* w2, int P, int H, int D, int dk, void* Wf,
                             void* Wd, void* stream);
int pka_head_grad_relayout(const float* dWcat, float* g0, float* g1, float* g2, int P, int H, int D, int dk,
                           void* stream);
/* pka_tc_reduce + pka_head_grad_relayout in one launch: g_p[h,d,j] = sum_{s<splits} ws[s][(p*H+h)*dk+j][d] (fixed order) */
int pka_tc_reduce_heads(const float* ws, float* g0, float* g1, float* g2, int splits, int P, int H, int D, int dk,
                        void* stream);

/* ---- batched helper launches of the training step (csrc/batched.cu) ---------------------------------------------
 * The reference's optimiser sees finished gradients because ATen finishes every reduction inside its op
 * (torch.autograd of Linear / LayerNormalization / bmm: L/train.py:197-199 calls backward() then step()).  Here the
 * backward kernels leave fixed-order split partials behind (weight-gradient splits, bias / LayerNorm column sums,
 * packed per-head gradients) and ONE launch per backward pass sums them all into the gradient arena:
 *   PLAIN: dst[e] (+)= sum_{s<splits} src[s*split_stride + e],  e < n
 *   HEADS: dst[(h*D + d)*dk + j] = sum_s src[s*split_stride + (h*dk + j)*D + d]   (one block [(h,j), d] of a packed
 *          head-projection gradient -> the reference's per-head layout [H, D, dk], T/SubLayers.py:29-31); n = H*D*dk.
 * Summation order is a fixed function of (splits), independent of the launch: bit-reproducible. */
enum { PKA_REDUCE_PLAIN = 0, PKA_REDUCE_HEADS = 1 };
#define PKA_MAX_REDUCE_JOBS 128
typedef struct {
  const float* src; float* dst;
  int64_t n, split_stride;
  int32_t splits, kind, accumulate, D, dk, reserved;
} pka_reduce_job;
int pka_reduce_jobs(const pka_reduce_job* jobs_host, int n_jobs, void* stream);
/* ONE launch refreshes the bf16 GEMM operand copies of all fp32 master weights (they change only in the optimiser
 * step; round 1 re-made them in every forward, 29 launches) and, optionally, advances the dropout step counter:
 *   PLAIN: src fp32 [N, nseg*K] -> wf bf16 [N, ldf] (same element order) and/or wd bf16 [K, ldd], wd[i, s*N+o] = src[o, s*K+i]
 *   HEADS: src fp32 [H=N, D=K, dk=nseg] -> rows n0 + h*dk + j of wf[.., ldf] (wf[n, d]) and columns n0 + h*dk + j of
 *          wd [D, ldd] (wd[d, n]): one block of a packed q|k|v operand (or of the k|v operand of several layers). */
enum { PKA_RELAYOUT_PLAIN = 0, PKA_RELAYOUT_HEADS = 1 };
#define PKA_MAX_RELAYOUT_JOBS 128
typedef struct {
  const float* src; void* wf; void* wd;
  int32_t N, K, nseg, kind, ldf, ldd, n0, reserved;
} pka_relayout_job;
int pka_relayout_jobs(const pka_relayout_job* jobs_host, int n_jobs, uint64_t* step_counter, void* stream);
/* teacher-forcing split of L/train.py:163-165 in one launch: tgt i64[B,L1], mask u8[B,L1] ->
 * tgt_in = tgt[:, :-1], goal = tgt[:, 1:], mask_in = mask[:, :-1]  (contiguous [B, L1-1]) */
/* y bf16 [rows, Np] = x [rows, N] (fp32 or bf16) with zero columns N..Np-1: makes the gradient of a [.., V = 53] output a
 * TMA-legal GEMM operand (16-byte row pitch), so the vocabulary projection's backward runs on the tensor cores */
int pka_pad_cast(const void* x, int dtype, void* y, int64_t rows, int N, int Np, void* stream);
int pka_split_targets(const int64_t* tgt, const uint8_t* mask, int64_t* tgt_in, int64_t* goal, uint8_t* mask_in, int B,
                      int L1, void* stream);

/* dZ = gate ? ((Y > 0) ? dY*scale : 0) : dY, bf16, written row-major [Bt*T, N] (dZ) and transposed [N, Bt, Tp] (dZt) */
int pka_relu_bwd_dual(const void* dY, int dy_dtype, const void* Y, void* dZ, void* dZt, int Bt, int T, int Tp, int N,
                      float scale, int gate, void* stream);

/* ---- (b) attention ----------------------------------------------------------------------------------------------
 * replaces: ScaledDotProductAttention.forward (T/Modules.py:75-97) with the masks of T/Models.py:27-49 evaluated
 * in-kernel.  q[B,Lq,ldq] k[B,Lk,ldk] v[B,Lk,ldv]: head h lives in columns [h*dk,(h+1)*dk).  key_mask u8[B,Lk]
 * (1 = real).  Allowed pairs: key real AND (no band OR i+band_start <= j <= i+band_end).  Rows without an allowed
 * key give 0 output and 0 gradient.  lse f32[B,H,Lq] (log-sum-exp of scaled scores; -inf for dead rows).
 * probs_out (optional, f32[B,H,Lq,Lk]) materialises the post-softmax, pre-dropout probabilities the reference
 * returns as `attn`. */
typedef struct {
  int32_t B, H, Lq, Lk, dk, dv;
  int32_t ldq, ldk, ldv, ldo;
  int32_t use_band, band_start, band_end;
  float scale;
  pka_dropout drop;
} pka_attn_desc;
int pka_attn_fwd(const pka_attn_desc* d, int dtype, const void* q, const void* k, const void* v,
                 const uint8_t* key_mask, void* out, float* lse, float* probs_out, void* stream);
/* dq/dk/dv use the same leading dimensions as q/k/v. delta_ws: float[B*H*Lq] scratch. */
int pka_attn_bwd(const pka_attn_desc* d, int dtype, const void* q, const void* k, const void* v,
                 const uint8_t* key_mask, const void* out, const void* dout, const float* lse, float* delta_ws,
                 void* dq, void* dk, void* dv, void* stream);

/* Tensor-core forward of the same op for bf16 q/k/v with dk = dv = 64 (tcgen05.mma S = Q K^T and O = P V with fp32
 * accumulators in TMEM, Q/K/V tiles staged by TMA, online softmax between the two MMAs; the mask stays a predicate).
 * out: bf16 or f32 [B,Lq,ldo] (out_dtype); lse as above.  The matching backward is pka_attn_bwd on the same buffers
 * (dtype PKA_BF16) -- both regenerate identical dropout bits. */
int pka_attn_tc_fwd(const pka_attn_desc* d, const void* q, const void* k, const void* v, const uint8_t* key_mask,
                    void* out, int out_dtype, float* lse, void* stream);

/* Tensor-core backward for the same bf16 buffers (q/k/v/out/dout and dq/dk/dv all bf16, leading dimensions as in `d`):
 * two launches of one kernel template -- dQ over query tiles, then dK/dV over key tiles -- each recomputing
 * S = Q K^T and dP = dO V^T with tcgen05.mma, forming dS = P (M dP - delta) in registers and accumulating
 * dQ = dS K / dK = dS^T Q / dV = (P M)^T dO in TMEM.  No atomics: bit-reproducible.  delta_ws: float[B*H*Lq]. */
int pka_attn_tc_bwd(const pka_attn_desc* d, const void* q, const void* k, const void* v, const uint8_t* key_mask,
                    const void* out, const void* dout, const float* lse, float* delta_ws, void* dq, void* dk, void* dv,
                    void* stream);

/* ---- (c) fused [dropout] + residual add + LayerNormalization ----------------------------------------------------
 * replaces: `layer_norm(dropout(x) + residual)` (T/SubLayers.py:65-68,85-86) with LayerNormalization of
 * T/Modules.py:42-51: y = (z-mean)/(std_unbiased+eps)*a + b.  The "identity when size(1)==1" rule is applied by the
 * caller (it is a shape rule).  mean/rinv f32[rows] are saved for backward (rinv = 1/(std+eps)). */
int pka_add_layernorm_fwd(const void* x, const void* residual, const float* a, const float* b, void* y, float* mean,
                          float* rinv, int dtype, int rows, int D, float eps, const pka_dropout* drop, void* stream);
/* dres always written; dx written only when drop->p > 0 (otherwise dx == dres and may be NULL).
 * dab_ws: float[3*D*pka_ln_bwd_blocks(rows)] scratch, layout [block][da | db | column sums of the dx written][D]
 * (the third plane is the bias gradient of a linear layer that produced x); da/db are overwritten.
 * da == db == NULL: the caller sums the pka_ln_bwd_blocks(rows) partial rows itself (pka_reduce_jobs). */
int pka_ln_bwd_blocks(int rows);
int pka_add_layernorm_bwd(const void* dy, const void* x, const void* residual, const float* a, const float* mean,
                          const float* rinv, void* dx, void* dres, float* da, float* db, float* dab_ws, int dtype,
                          int rows, int D, float eps, const pka_dropout* drop, void* stream);

/* ---- (d) summed cross-entropy with PAD mask, optional label smoothing, fused argmax accuracy -------------------
 * replaces: cal_loss / get_performance (L/train.py:58-90).  logits[N,V] (ld = V), goal i64[N], PAD = 0 ignored.
 * out3: float[3] = {loss_sum, n_correct, n_words}; lse f32[N] saved for backward; part_ws float[3*pka_ce_blocks(N)]. */
int pka_ce_blocks(int N);
int pka_ce_fwd(const void* logits, const int64_t* goal, int dtype, int N, int V, int smoothing, float eps,
               float* out3, float* lse, float* part_ws, void* stream);
/* the same in ONE launch: the last CTA to finish sums the partials in the same fixed order.  done_counter: uint32[1],
 * zero before the first call (it resets itself). */
int pka_ce_fwd_fused(const void* logits, const int64_t* goal, int dtype, int N, int V, int smoothing, float eps,
                     float* out3, float* lse, float* part_ws, uint32_t* done_counter, void* stream);
/* dlogits = grad_out[0] * (softmax - target) on non-PAD rows, 0 on PAD rows. */
int pka_ce_bwd(const void* logits, const int64_t* goal, const float* lse, const float* grad_out, void* dlogits,
               int dtype, int N, int V, int smoothing, float eps, void* stream);

/* ---- element-wise pieces of the path ----------------------------------------------------------------------------
 * out[b,l,:] = dropout(emb[tok[b,l],:] + pos[l,:])           (T/Models.py:195-213) */
int pka_embed_pos_fwd(const int64_t* tok, const float* emb, const float* pos, void* out, int dtype, int B, int L,
                      int D, int V, const pka_dropout* drop, void* stream);
/* demb[v,:] = sum over (b,l) with tok==v of dropout_bwd(dout[b,l,:]); every row is written (row `padding_idx` and
 * rows of unseen tokens get zeros), so the caller can pass the gradient slot itself without clearing it.
 * Deterministic (one CTA per vocabulary row, fixed summation order). */
int pka_embed_bwd(const int64_t* tok, const void* dout, float* demb, int dtype, int B, int L, int D, int V,
                  int padding_idx, const pka_dropout* drop, void* stream);
/* out[r,:] = dropout(x[r,:] + rowvec[(r % period),:])  (rowvec may be NULL)   (T/Models.py:164-165,226) */
int pka_add_rowvec_dropout_fwd(const void* x, const float* rowvec, void* out, int dtype, int64_t rows, int D,
                               int period, const pka_dropout* drop, void* stream);
/* dx = keep * dy / (1-p) */
int pka_dropout_bwd(const void* dy, void* dx, int dtype, int64_t n, const pka_dropout* drop, void* stream);
/* dz = (y > 0) ? dy * scale : 0  -- backward of ReLU followed by dropout, using only the layer output y
 * (y > 0  <=>  pre-activation > 0 and kept).  scale = 1/(1-p). */
int pka_relu_drop_bwd(const void* dy, const void* y, void* dz, int dtype, int64_t n, float scale, void* stream);
/* out[n] (+)= sum_r x[r*ld + n]; part_ws float[pka_colsum_chunks(rows)*N]; deterministic two-stage reduction.
 * out == NULL (here and in pka_gate_colsum): only the partial rows are written, the caller sums the
 * pka_colsum_chunks(rows) rows of part_ws itself (pka_reduce_jobs). */
int pka_colsum_chunks(int64_t rows);
/* partial rows actually written for these arguments (<= pka_colsum_chunks(rows)) */
int pka_colsum_parts(const void* x, const float* part_ws, int dtype, int64_t rows, int N, int ld);
int pka_gate_colsum_parts(int dy_dtype, int64_t rows, int N);
int pka_colsum(const void* x, float* out, float* part_ws, int dtype, int64_t rows, int N, int ld, int accumulate,
               void* stream);
/* one pass: dZ (bf16 [rows,N]) = gate ? ((Y > 0) ? dY*scale : 0) : dY, and out[n] = sum_r dZ[r,n] (bias gradient of a
 * [ReLU+dropout] linear layer).  part_ws float[pka_colsum_chunks(rows)*N]; deterministic.  N % 4 == 0. */
int pka_gate_colsum(const void* dY, int dy_dtype, const void* Y, void* dZ, float* out, float* part_ws, int64_t rows, int N,
                    float scale, int gate, void* stream);
/* keep[i] = 1/0 for i < n : the bits every kernel above derives for (seed, site, *step_ptr) */
int pka_dropout_mask(uint8_t* keep, int64_t n, const pka_dropout* drop, void* stream);
int pka_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);
/* dst[c, r] = src[r, c] for a [rows, cols] matrix (dtype conversion allowed) */
int pka_transpose(const void* src, int src_dtype, void* dst, int dst_dtype, int rows, int cols, void* stream);

/* ---- optimiser ---------------------------------------------------------------------------------------------------
 * replaces: torch.optim.Adam.step over 72 tensors + ScheduledOptim.update_learning_rate (T/Optim.py:13-27,
 * L/train.py:376-380) with one launch over the flat parameter arena.
 * state: int64[2] = {adam_t, n_current_steps}; lr: float[1] on device.  pka_adam_step uses t = adam_t+1 and then
 * stores it; pka_lr_tick does n+=1; lr = start_lr*c/(n+c).  bf16_shadow (optional) receives the updated weights
 * rounded to bf16 for the tensor-core path.  If grad_scale_inv != NULL gradients are multiplied by it first. */
int pka_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr_dev,
                  float lr_host, int64_t* state, float beta1, float beta2, float eps, void* bf16_shadow,
                  void* stream);
int pka_lr_tick(float* lr_dev, int64_t* state, float start_lr, float soft_coefficient, void* stream);
/* pka_adam_step + the counter updates in ONE launch: the last CTA to finish stores adam_t+1 and, if tick_lr, performs
 * pka_lr_tick's update (every CTA has read adam_t / lr before any CTA can be last).  done_counter: uint32[1], zero
 * before the first call (it resets itself). */
int pka_adam_step_fused(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float* lr_dev,
                        float lr_host, int64_t* state, float beta1, float beta2, float eps, void* bf16_shadow,
                        uint32_t* done_counter, int tick_lr, float start_lr, float soft_coefficient, void* stream);

/* ---- host side of the input feed (no device work) --------------------------------------------------------------------
 * replaces: pad_to_longest over the features of a batch (U/instances_handler.py:118-139, U/BatchLoader.py:92-103).
 * Gathers n_utt ragged utterances (numbers idx[]) from a packed frame store frames[total, dim] / offsets[n_total + 1]
 * into src[n_utt, t_pad, dim] (zero padded) and mask[n_utt, t_pad] (1 = real frame), with up to n_threads threads.
 * All pointers are HOST memory; src / mask are normally the pinned staging buffers of the H2D copy. */
int pka_host_pack_batch(const float* frames, const int64_t* offsets, int64_t n_total, const int64_t* idx, int n_utt,
                        int t_pad, int dim, float* src, uint8_t* mask, int n_threads);

/* ---- data-parallel optimiser step over peer memory (NVLink / NVSwitch) ------------------------------------------
 * replaces, for W ranks: {all-reduce SUM of the gradient arena (the new exchange step of SURVEY.md 8e), pka_adam_step}
 * by ONE kernel per rank: gradient reduce-scatter (multimem.ld_reduce through the switch when multicast addresses
 * are given, W peer loads in rank order otherwise), Adam on the rank's 1/W shard of (param, exp_avg, exp_avg_sq),
 * parameter all-gather (multimem.st / W peer stores), bracketed by two per-CTA flag handshakes with the peers.
 *   param_ptrs / grad_ptrs : HOST arrays of W device addresses -- the symmetric parameter / gradient arena as mapped
 *                            for rank 0..W-1 (entry `rank` is the local one); n floats each, n % (4*W) == 0
 *   param_mc / grad_mc     : NVSwitch multicast addresses of the two arenas, or 0 / 0
 *   flag_ptrs_dev          : DEVICE array of W device addresses of the ranks' flag arrays, each
 *                            uint32[2 * pka_dp_adam_grid(n, W, max_ctas) * W], zero before the first call
 *   exp_avg / exp_avg_sq   : local, n floats; only the rank's shard [rank*n/W, (rank+1)*n/W) is read and written
 *   state, lr_dev, lr_host, beta1, beta2, eps: as pka_adam_step (adam_t is advanced on every rank)
 * All ranks must call it with the same n, W and max_ctas, once per step, in the same order relative to each other.
 * CUDA-graph capturable; the grid (<= 148 CTAs) has to be co-resident with whatever else runs on the device. */
int pka_dp_adam_grid(int64_t n, int world, int max_ctas);
int pka_dp_adam_step(const uint64_t* param_ptrs, const uint64_t* grad_ptrs, uint64_t param_mc, uint64_t grad_mc,
                     const uint64_t* flag_ptrs_dev, int rank, int world, int max_ctas, float* exp_avg, float* exp_avg_sq,
                     int64_t n, const float* lr_dev, float lr_host, int64_t* state, float beta1, float beta2, float eps,
                     void* stream);
int pka_counter_inc(uint64_t* counter, void* stream);

/* ---- (e) beam-search decoding ------------------------------------------------------------------------------------
 * replaces: the per-step body of translate_batch (L/decode.py:54-98) + Lattice.advance (T/Lattice.py:35-81).
 * Device-resident lattice, one row per utterance, E = max_edges >= 1 + beam*max_len (edge 0 = BOS):
 *   edge_prev/edge_word/edge_depth int32[n_utt,E], edge_weight f64[n_utt,E], n_edges int32[n_utt]
 *   beam_edges int32[n_utt,beam] + beam_count int32[n_utt]: current beam, best first (finished edges stay in it)
 *   slot_edge/slot_active int32[n_utt,beam]: live hypotheses in beam order = rows (u*beam+slot) of the decoder step
 *   curr_length/done int32[n_utt], n_not_done int32[1] */
typedef struct {
  int32_t n_utt, beam, V, max_edges, max_len;
  int32_t eos, force_full_length;
  int32_t inputs_are_logprobs;   /* 1: `logits` already holds log-probabilities (Lattice.advance API); skip the log-softmax */
} pka_beam_desc;
/* fp32 log-softmax of logits[n_utt*beam, V] rows of live slots; candidate = parent weight (f64) + log-prob; finished
 * hypotheses appended; warp-shuffle top-`beam` (ties -> lowest flat index); lattice and slots rewritten in place.
 * force_full_length: the EOS log-prob is set to -1e30 (benchmark stress variant: nobody ever finishes). */
int pka_beam_advance(const pka_beam_desc* d, const float* logits, int32_t* edge_prev, int32_t* edge_word,
                     int32_t* edge_depth, double* edge_weight, int32_t* n_edges, int32_t* beam_edges,
                     int32_t* beam_count, int32_t* slot_edge, int32_t* slot_active, int32_t* curr_length,
                     int32_t* done, int32_t* n_not_done, void* stream);
/* self-attention of the newest token of every live slot over itself (k_self/v_self rows of the packed qkv buffer,
 * leading dimension ld) and its <= window-1 ancestors, read from the per-edge caches kcache/vcache
 * f32[n_utt, E, H*dk] along edge_prev.  out f32[n_utt*beam, H*dk]; dead slots give zeros. */
int pka_tree_attn(const float* q, const float* k_self, const float* v_self, int ld, const float* kcache,
                  const float* vcache, const int32_t* edge_prev, const int32_t* slot_edge, const int32_t* slot_active,
                  float* out, int n_utt, int beam, int max_edges, int H, int dk, int window, float scale, void* stream);
/* scatter the newest token's K/V rows (slot-major, leading dimension ld) into the per-edge caches */
int pka_kv_append(const float* k_new, const float* v_new, int ld, float* kcache, float* vcache,
                  const int32_t* slot_edge, const int32_t* slot_active, int n_utt, int beam, int max_edges, int HD,
                  void* stream);
/* decoder input of every slot: emb[word(edge)] + pos[depth(edge)]; dead slots give zeros */
int pka_beam_embed(const float* emb, const float* pos, const int32_t* edge_word, const int32_t* edge_depth,
                   const int32_t* slot_edge, const int32_t* slot_active, float* out, int n_utt, int beam,
                   int max_edges, int D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PKA_B200_H */
