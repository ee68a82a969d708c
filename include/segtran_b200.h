/*
 * segtran_b200 — C ABI of the B200-native Squeeze-and-Expansion hot path.
 *
 * The reference (askerlee/segtran) is pure Python/PyTorch and has no FFI layer: the drop-in
 * boundary is the nn.Module contract of code/networks/segtran_shared.py (SegtranFusionEncoder
 * :819-975 and the modules it owns) and of the Segtran2d/Segtran3d shells.  This header is the
 * C ABI introduced *underneath* that contract: every entry point names the
 * reference call sites whose arithmetic it replaces.  Host binding: ctypes (segtran_b200/_lib.py);
 * see INTEGRATION.md for the binding a maintainer of the reference would add.
 *
 * Conventions
 *  - every pointer is a caller-owned DEVICE pointer (tensor.data_ptr()); nothing is retained
 *    after the call returns; scratch is passed in explicitly by the caller;
 *  - every call is asynchronous on the cudaStream_t given (passed as void*);
 *  - return 0 on success, negative on error; sx_last_error() returns a thread-local message;
 *  - no global mutable state apart from one-time function-attribute setup (thread safe: the
 *    autograd engine calls backward entry points from its own worker thread);
 *  - the device is the current CUDA context's device (torch sets it); never assumed to be 0.
 *  - there is NO CPU fallback: a missing GPU / non-sm_100 device is an error.
 */
#ifndef SEGTRAN_B200_H_
#define SEGTRAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SX_VERSION 1

/* element types */
enum { SX_F32 = 0, SX_BF16 = 1 };
/* GEMM operand arithmetic: TF32 (fp32 storage, 10-bit mantissa in the tensor core) or BF16 */
enum { SX_OP_TF32 = 0, SX_OP_BF16 = 1 };
/* operand majorness: K-major = reduction dim contiguous; MN-major = row/col dim contiguous */
enum { SX_MAJOR_K = 0, SX_MAJOR_MN = 1 };
enum { SX_BIAS_NONE = 0, SX_BIAS_N = 1, SX_BIAS_M = 2 };
enum { SX_ACT_NONE = 0, SX_ACT_GELU = 1,
       SX_ACT_GELU_BWD = 2 /* C = dropmask * (alpha A.B^T) * gelu'(preact): `preact` is an INPUT in C's layout */ };

int sx_version(void);
const char* sx_last_error(void);
/* number of SMs / compute capability of the current device (diagnostics; fails if no GPU) */
int sx_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * Batched GEMM on tcgen05 tensor cores (TMA-staged operands, fp32 accumulators in TMEM):
 *     C[z1][z0][m][n] = epilogue( alpha * sum_k A[z1][z0][m][k] * B[z1][z0][n][k] )
 * Replaces every dense contraction on the path: the Q/K/V projections (segtran_shared.py:559-560,
 * :414), Q.K^T (:566), P.V (:447), MMSharedMid's shared Linear (:243), MMPrivateOutput's grouped
 * Conv1d (:267), and all of their backward products.
 * An operand with majorness K stores element (r,k) at ptr[r*ld + k]; majorness MN at ptr[k*ld + r].
 * stride_z0/stride_z1 are in elements; 0 broadcasts the operand over that batch dim.
 * Requirements: ptr 16-byte aligned; ld*elsize and z strides*elsize multiples of 16 bytes.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* ptr;
  int32_t major;
  int32_t _pad;
  int64_t ld;
  int64_t stride_z0;
  int64_t stride_z1;
} sx_operand;

typedef struct {
  int32_t op_dtype;              /* SX_OP_TF32 | SX_OP_BF16 (A and B element type: f32 | bf16) */
  int32_t M, N, K, Z0, Z1;
  sx_operand A, B;
  void* C;
  int32_t c_dtype;               /* SX_F32 | SX_BF16 */
  int32_t round_tf32;            /* round fp32 outputs to TF32 (RN) so a following TF32 GEMM is exact on them */
  int64_t ldc, c_stride_z0, c_stride_z1;
  float alpha;
  int32_t bias_mode;             /* SX_BIAS_* ; bias is fp32 */
  const float* bias;
  int64_t bias_stride_z0, bias_stride_z1;
  int32_t act;                   /* SX_ACT_* */
  int32_t accumulate;            /* 0: store, 1: atomicAdd into fp32 C (needed when split_k > 1) */
  void* preact;                  /* optional: pre-activation (after bias) in C's layout and dtype */
  int32_t split_k;               /* >= 1 */
  int32_t _pad2;
  float* amax;                   /* optional: atomicMax of the stored values (attention-score diagnostics, :569-573) */
  float drop_p;                  /* dropout on the stored value (after act); 0 disables */
  uint32_t _pad3;
  uint64_t drop_seed;            /* counter-based mask: keep(idx) = hash(seed, flat index in C) >= p */
  const uint64_t* drop_seed_dev; /* optional device seed added to drop_seed (CUDA-graph safe) */
  const float* addend;           /* optional fp32 tensor in C's layout: C = epilogue(alpha*A.B^T + addend); used by the
                                    error-compensated 3-pass TF32 mode (A_hi B_hi + A_lo B_hi + A_hi B_lo) */
  float* colsum;                 /* optional [N] fp32: += column sums of the stored values over all rows and batch slices
                                    (a bias gradient that would otherwise need its own pass over the output) */
} sx_gemm_args;

int sx_gemm(const sx_gemm_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused attention probabilities of the squeeze-out stage (csrc/sx_attn.cu):
 *     P[b][m] = dropout( softmax_keys( min(alpha * Q[b,:,m] K[b,:,m]^T, clip) ) )
 * One persistent tcgen05 kernel replaces segtran_shared.py:566-567 (Q.K^T / sqrt(d)), :569-580 (max statistics and
 * conditional clamp), :601 (softmax) and :605 (attention dropout): the scores stay in TMEM, the softmax runs on the
 * tcgen05.ld fragments, only P is written (plus the raw scaled scores S when the backward needs them).  More than 256
 * keys do not fit one TMEM accumulator per row block: the scores are then produced twice (statistics launch, then
 * probabilities launch, both over (row block, key chunk) tiles) instead of being written and re-read.
 * Q [Bq][U1][M*d] (q_bstride = 0: one query bank shared by the batch), K [B][U2][M*d]; mode m uses columns
 * [m*d, (m+1)*d).  P, S: [B][M][U1][ldp] fp32, ldp % 4 == 0.  lse, rowmax: [B][M][U1] (natural-log units; rowmax is
 * the max of the raw row).  stat: device scratch of three 32-bit words, ZERO-initialised by the caller: [0] the running
 * maximum of the scores under an order-preserving float->uint map (0 = none yet), [1] (float) number of rows whose
 * maximum is below -(clip - 104), [2] (written at the end) the maximum as a plain float — the `amax` of sx_softmax_bwd.  diag (optional, device float[3]): [0] running
 * max, [1] += 1 when the clamp fired (max > clip), [2] += stat[1] in that case (rows where the reference's LOWER
 * clamp could have mattered — the upper clamp is applied exactly; see sx_attn.cu).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, M, U1, U2, d;
  int32_t round_tf32;
  const float* Q;
  int64_t q_ld, q_bstride;
  const float* K;
  int64_t k_ld, k_bstride;
  float alpha, clip;
  float* P;
  float* S;                      /* optional */
  int64_t ldp;
  float* lse;
  float* rowmax;                 /* optional */
  float* stat;
  float* diag;                   /* optional */
  float drop_p;
  uint32_t _pad;
  uint64_t drop_seed;
  const uint64_t* drop_seed_dev;
  float* scratch;                /* U2 > 256 only: B*M*ceil(U1/256)*ceil(U2/256)*1536 floats of partial row statistics */
  int64_t scratch_floats;
} sx_attn_probs_args;
int sx_attn_probs_fwd(const sx_attn_probs_args* args, void* stream);

/* Dropout seeds: every dropout-capable entry takes `seed` (by value) and `seed_dev` (device pointer or NULL); the
 * effective seed is seed + *seed_dev, read on the device in stream order, so a captured CUDA graph draws a new mask
 * on every replay.  sx_seed_derive writes out[0] = base[0] + add (the per-call seed an op keeps for its backward);
 * sx_seed_advance bumps the base seed once per training step. */
int sx_seed_derive(const uint64_t* base, uint64_t add, uint64_t* out, void* stream);
int sx_seed_advance(uint64_t* base, uint64_t inc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Row-wise kernels (HBM-bound).  "act dtype" arguments are SX_F32 | SX_BF16; round_tf32 rounds fp32
 * outputs that feed a TF32 GEMM.  Dropout masks are counter-based: keep(i) = hash(seed, flat index)
 * >= p, so backward regenerates the forward mask from (seed, index) instead of storing it.
 * ------------------------------------------------------------------------------------------- */

/* out[0] = max(x[0..n))                        -- voxels_pos.max(), segtran_shared.py:1231 */
int sx_reduce_max(const float* x, int64_t n, float* out, void* stream);

/* Learnable-sinusoid positional code, LearnedSinuPosEmbedder.forward (segtran_shared.py:989-998) with the
 * pos/pos.max() normalisation of SegtranPosEncoder.forward (:1231).  pos [R,pd], W [C,pd], b [C] -> pe [R,C]. */
int sx_pos_lsinu_fwd(const float* pos, const float* posmax, int64_t R, int32_t pd, const float* W, const float* b,
                     int32_t C, float* pe, void* stream);
/* dpe [R,C] -> dW [C,pd], db [C] accumulated (+=); de_scratch [R,C] is caller-provided scratch. */
int sx_pos_lsinu_bwd(const float* pos, const float* posmax, int64_t R, int32_t pd, const float* W, const float* b,
                     int32_t C, const float* dpe, float* de_scratch, float* dW, float* db, void* stream);

/* Fused prologue of SegtranFusionEncoder.forward (segtran_shared.py:916, :930-934, :944-946):
 *   h = mask * dropout( LN( LN_{g,b}(x) + posw * pe[..., :C] ) ),  x [B,N,C] fp32, pe rows of length C0,
 *   pe_bstride = 0 when the code is shared by the batch; mask [B*N] fp32 or NULL; stats [B*N,4]. */
int sx_prologue_fwd(const float* x, int64_t B, int32_t N, int32_t C, const float* g, const float* b, const float* pe,
                    int32_t C0, int64_t pe_bstride, float posw, const float* mask, float drop_p, uint64_t seed, const uint64_t* seed_dev, void* h,
                    int32_t h_dtype, int32_t round_tf32, float* stats, void* stream);
/* dh fp32 -> dx [B,N,C]; dg, db [C] and dpe (same addressing as pe, may be NULL) are accumulated (+=). */
int sx_prologue_bwd(const float* dh, const float* x, int64_t B, int32_t N, int32_t C, const float* g, const float* b,
                    const float* pe, int32_t C0, int64_t pe_bstride, float posw, const float* mask, float drop_p,
                    uint64_t seed, const uint64_t* seed_dev, const float* stats, float* dx, float* dg, float* db, float* dpe,
                    float* dt_scratch /* [B*N*C] or NULL */, void* stream);

/* Row softmax with the reference's conditional clamp and attention dropout (segtran_shared.py:578-580, :601-605):
 *   if (*amax > clip) S = clamp(S, -clip, clip);  P = dropout(softmax(S)).  S [R,L] fp32 (row stride lds),
 *   P [R,L] (row stride ldp), lse [R] = log-sum-exp of the (clamped) row, kept for backward.
 *   diag (optional, device float[2]): [0] = running max of *amax, [1] += 1 when the clamp fired — the module's
 *   max_attn / clamp_count counters (:575-587) without the reference's two .item() host syncs per call. */
int sx_softmax_fwd(const float* S, int64_t R, int32_t L, int64_t lds, const float* amax, float clip, float drop_p,
                   uint64_t seed, const uint64_t* seed_dev, void* P, int32_t p_dtype, int64_t ldp, int32_t round_tf32, float* lse, float* diag,
                   void* stream);
int sx_softmax_bwd(const float* dP, int64_t ldd, const float* S, int64_t lds, const float* lse, int64_t R, int32_t L,
                   const float* amax, float clip, float drop_p, uint64_t seed, const uint64_t* seed_dev, int64_t ldp_fwd, void* dS,
                   int32_t ds_dtype, int64_t ldo, int32_t round_tf32, void* stream);

/* LayerNorm with affine over rows, eps 1e-12 (first_norm_layer, segtran_shared.py:456).  stats [R,2]. */
int sx_layernorm_fwd(const float* x, int64_t R, int32_t C, const float* g, const float* b, void* y, int32_t y_dtype,
                     int32_t round_tf32, float* stats, void* stream);
int sx_layernorm_bwd(const float* dy, const float* x, int64_t R, int32_t C, const float* g, const float* stats,
                     void* dx, int32_t dx_dtype, int32_t round_tf32, float* dg, float* db, void* stream);

/* MMPrivateOutput tail + LearnedSoftAggregate (segtran_shared.py:273-274, :318-325):
 *   Yn = LN_{g,b}(dropout(Y));  w = softmax_modes(Yn.ws + bs);  out = sum_m w_m Yn_m
 *   Y [B,M,N,F] fp32 -> out [B,N,F] fp32; stats [B,M,N,2]; wts [B,M,N]. */
int sx_ln_softaggr_fwd(const float* Y, int32_t B, int32_t M, int32_t N, int32_t F, const float* g, const float* b,
                       const float* ws, const float* bs, float drop_p, uint64_t seed, const uint64_t* seed_dev, float* out, float* stats,
                       float* wts, void* stream);
int sx_ln_softaggr_bwd(const float* dout, const float* Y, int32_t B, int32_t M, int32_t N, int32_t F, const float* g,
                       const float* b, const float* ws, float drop_p, uint64_t seed, const uint64_t* seed_dev, const float* stats,
                       const float* wts, void* dY, int32_t dy_dtype, int32_t round_tf32, float* dg, float* db,
                       float* dws, float* dbs, float* dscore_scratch /* [B*M*N] or NULL */, void* stream);

/* LearnedSoftAggregate on its own (segtran_shared.py:318-325; the no-FFN branch :453 with M modes — the Polyformer layer):
 *   w = softmax_modes(x_m . ws + bs);  out = sum_m w_m x_m.   x [B,M,N,F] -> out [B,N,F], wts [B,M,N].
 * backward: dx [B,M,N,F] and dscore [B,M,N] (d ws = sum dscore x, d bs = sum dscore are left to the caller). */
int sx_softaggr_fwd(const float* x, int32_t B, int32_t M, int32_t N, int32_t F, const float* ws, const float* bs, float* out,
                    float* wts, void* stream);
int sx_softaggr_bwd(const float* dout, const float* x, int32_t B, int32_t M, int32_t N, int32_t F, const float* ws,
                    const float* wts, float* dx, float* dscore, void* stream);
/* dH = dropout'(dG) * gelu'(H)  (MMSharedMid backward, segtran_shared.py:243-245) */
int sx_gelu_bwd(const float* dG, const void* H, int32_t h_dtype, int64_t n, float drop_p, uint64_t seed, const uint64_t* seed_dev, void* dH,
                int32_t dh_dtype, int32_t round_tf32, void* stream);
/* dtype conversion / TF32 rounding of a flat buffer (weights once per step) */
int sx_convert(const void* x, int32_t x_dtype, int64_t n, void* y, int32_t y_dtype, int32_t round_tf32, void* stream);
/* hi = TF32(x), lo = TF32(x - hi) over a flat fp32 buffer: operand split of the 3-pass error-compensated TF32 products
 * (A_hi B_hi + A_lo B_hi + A_hi B_lo) used by the precision policy for the small / sensitive contractions */
int sx_split_tf32(const float* x, int64_t n, float* hi, float* lo, void* stream);
/* x [Z1][Z0][R][K] with element strides (sz1, sz0, sr, sk) -> out [Z1][Z0][R][3*Kp] contiguous, rows = [lo|hi|hi] (role 0,
 * the "A" operand) or [hi|lo|hi] (role 1, the "B" operand), segments zero-padded to Kp (multiple of 4) columns: ONE
 * sx_gemm launch over K' = 3*Kp on the two outputs is the 3-pass product A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T */
int sx_split_tf32_cat(const float* x, int32_t Z1, int32_t Z0, int32_t R, int32_t K, int64_t sz1, int64_t sz0, int64_t sr,
                      int64_t sk, int32_t Kp, int32_t role, float* out, void* stream);
/* out[c] += sum_r X[r,c]   (bias gradients) */
int sx_colsum(const void* X, int32_t x_dtype, int64_t R, int32_t C, int64_t ld, float* out, void* stream);
/* out[0] += sum_i x[i]*y[i]  and  y = alpha * (*alpha_dev) * x : a linear loss head for benchmarks / checksums */
int sx_dot(const float* x, const float* y, int64_t n, float* out, void* stream);
int sx_scale(const float* x, int64_t n, const float* alpha_dev, float alpha, float* y, void* stream);
/* out[z0][c] += sum_{z1,r} X[z1][z0][r][c]  (per-mode bias gradient of MMPrivateOutput in one launch) */
int sx_colsum_batched(const float* X, int32_t Z1, int64_t stride_z1, int32_t Z0, int64_t stride_z0, int64_t R, int32_t C,
                      int64_t ld, float* out, void* stream);
/* y = a + b  (residual connection of MMSharedOutput, segtran_shared.py:305) */
int sx_add(const float* a, const float* b, int64_t n, float* y, void* stream);
/* out[r % out_mod] += sum_c X[r,c]  (class-bias gradient of the head: rows = (batch, class)) */
int sx_rowsum(const float* X, int64_t R, int64_t C, int64_t ld, int32_t out_mod, float* out, void* stream);
/* batched transpose [Z,R,C] -> [Z,C,R] fp32: token flatten / scatter (segtran3d.py:328-330, :478-480) */
int sx_transpose(const float* in, int64_t Z, int32_t R, int32_t C, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Segmentation head, collapsed form (segtran3d.py:364-367, :381-386, :488-496; segtran2d.py:304-306,
 * :427, :435-436).  curr [B,Cf,V] fp32 channels-first, W [K,Cf], L / dL [B,K,V].
 * ------------------------------------------------------------------------------------------- */
int sx_head_contract_fwd(const float* curr, const float* W, const float* bias, int32_t B, int32_t Cf, int64_t V,
                         int32_t K, float* L, int32_t accumulate, void* stream);
int sx_head_contract_bwd_data(const float* dL, const float* W, int32_t B, int32_t Cf, int64_t V, int32_t K,
                              float* dcurr, void* stream);
int sx_head_contract_bwd_weight(const float* dL, const float* curr, int32_t B, int32_t Cf, int64_t V, int32_t K,
                                float* dW, void* stream);
/* class scores of the fused tokens, exact fp32: out[b,k,n] = sum_f W[k,f] vf[b,n,f]   (Wc . vfeat_fused) */
int sx_token_scores(const float* vf, const float* W, int32_t B, int32_t N, int32_t F, int32_t K, float* out,
                    void* stream);
/* and its data gradient: dvf[b,n,f] = sum_k dt[b,k,n] W[k,f]   (F % 4 == 0) */
int sx_token_scores_bwd(const float* dt, const float* W, int32_t B, int32_t N, int32_t F, int32_t K, float* dvf,
                        void* stream);
/* -------------------------------------------------------------------------------------------
 * FPN pyramid stage (segtran3d.py:299-313, :347-359, segtran2d.py:244-300):
 *   curr <- GroupNorm_G( conv1x1(curr) + bias + upsample(higher) )
 * conv1x1 + bias + add is ONE sx_gemm launch on the channels-first tensors (A = W [Cout x Cin] broadcast over the batch,
 * B = x[b] read as an MN-major [V x Cin] operand, bias mode SX_BIAS_M, addend = the upsampled level); the upsampling is
 * sx_resize_axis_fwd per axis; GroupNorm is below.  x, y, dy, dx: [B, C, V] fp32.  csum: [B*C*2] double workspace,
 * stats: [B*G*2] (mean, rstd) kept for the backward, coef: [B*G*2] workspace.  dgamma/dbeta are accumulated into.
 * ------------------------------------------------------------------------------------------- */
int sx_groupnorm_fwd(const float* x, int32_t B, int32_t C, int64_t V, int32_t G, const float* gamma, const float* beta,
                     float eps, double* csum, float* stats, float* y, int32_t round_tf32, void* stream);
int sx_groupnorm_bwd(const float* dy, const float* x, int32_t B, int32_t C, int64_t V, int32_t G, const float* gamma,
                     const float* stats, double* csum, float* coef, float* dx, float* dgamma, float* dbeta, void* stream);

/* -------------------------------------------------------------------------------------------
 * Training-step tail: segmentation loss and BertAdam on flat buckets.
 * Everything the step needs (loss scalars, clip coefficients, scheduled learning rates, the step
 * counter) is produced and consumed on the device, so the step can be captured in a CUDA graph.
 * ------------------------------------------------------------------------------------------- */
/* loss = (1-dice_w) * BCEWithLogits(pos_weight)(logits, mask) + dice_w * sum_{k>=1} class_w[k] * dice_loss_indiv(sigmoid(logits[:,k]), mask[:,k])
 * (train3d.py:731-756, utils/losses.py:47-60).  logits, mask: [B,K,V] fp32 (mask n-hot).  pos_weight, class_w: [K] or NULL (= ones).
 * sums: [B*K*4] double workspace (zeroed here); out3 = {loss, ce, dice}; coef: [B*K*2] Dice gradient coefficients for the backward. */
int sx_seg_loss_fwd(const float* logits, const float* mask, int32_t B, int32_t K, int64_t V, const float* pos_weight,
                    const float* class_w, float dice_w, double* sums, float* out3, float* coef, void* stream);
/* dlogits = (*gout or 1) * d loss / d logits; ce_scale = (1-dice_w) / (B*K*V); coef from sx_seg_loss_fwd */
int sx_seg_loss_bwd(const float* logits, const float* mask, int32_t B, int32_t K, int64_t V, const float* pos_weight,
                    const float* coef, float ce_scale, const float* gout, float* dlogits, void* stream);
/* One optimiser step of the reference's BertAdam (optimization.py:90-164) preceded by the global gradient-norm clip of
 * train3d.py:760-761, on flat fp32 buckets p, g, m, v.  The buckets are cut into segments (<= a few thousand elements of
 * ONE parameter each): seg_param / seg_off / seg_len [nseg].  lr, wd: per-parameter [P].  schedule: SX_SCHED_*; t_total = -1
 * disables the schedule.  step: device counter (read, then incremented).  sumsq [P] double, coef [P], lr_eff [P]: device
 * workspaces.  total_norm: optional device float (the pre-clip global norm).  A parameter whose gradient is exactly zero is
 * left untouched (the reference skips p.grad is None: never-used parameters).  p_tf32 (optional): a second parameter
 * buffer that receives the TF32-rounded new values, so the next forward needs no per-weight rounding pass. */
enum { SX_SCHED_WARMUP_LINEAR = 0, SX_SCHED_WARMUP_CONSTANT = 1 };
int sx_adam_step(float* p, float* p_tf32, const float* g, float* m, float* v, const int32_t* seg_param, const int64_t* seg_off,
                 const int32_t* seg_len, int32_t nseg, int32_t P, const float* lr, const float* wd, double b1, double b2,
                 double eps, float grad_clip, float max_grad_norm, float warmup, int64_t t_total, int32_t schedule,
                 int64_t* step, double* sumsq, float* coef, float* lr_eff, float* total_norm, void* stream);

/* 1-D linear resampling (align_corners=False) of x viewed as [outer, Lin, inner] -> [outer, Lout, inner];
 * F.interpolate(mode='bilinear'|'trilinear') == one pass per axis. */
int sx_resize_axis_fwd(const float* x, int64_t outer, int32_t Lin, int32_t Lout, int64_t inner, float* y,
                       int32_t accumulate, void* stream);
int sx_resize_axis_bwd(const float* dy, int64_t outer, int32_t Lin, int32_t Lout, int64_t inner, float* dx,
                       void* stream);
/* tiny strided fp32 GEMM on CUDA cores (class-dimension products of the collapsed head):
 *   C[z](m,n) (+)= alpha * sum_k A[z](m,k) B[z](k,n), element strides given explicitly */
int sx_sgemm_small(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, int64_t sam, int64_t sak,
                   int64_t sbk, int64_t sbn, int64_t scm, int64_t scn, int32_t Z, int64_t saz, int64_t sbz, int64_t scz,
                   float alpha, int32_t accumulate, void* stream);

/* -------------------------------------------------------------------------------------------
 * Sliding-window inference post-process (code/test_util3d.py:93-184):
 * sx_sw_accumulate: preds[k][window] += sigmoid(scores[k]), cnt[window] += 1 for one patch ([K][dx][dy][dz] scores, window
 *   origin (x0,y0,z0) in the [K][H][W][D] accumulators)                                              (test_util3d.py:155-159)
 * sx_sw_finalize: preds /= cnt; brats: make_brats_pred_consistent(is_conservative=False) (datasets3d.py:53-59), hard[1:] =
 *   preds >= 0.5, hard[0] = no class fired (hard is [K][V]); otherwise hard[0..V) = argmax_k as a float class index.
 * ------------------------------------------------------------------------------------------- */
int sx_sw_accumulate(const float* scores, int32_t K, int32_t dx, int32_t dy, int32_t dz, float* preds, float* cnt,
                     int32_t H, int32_t W, int32_t D, int32_t x0, int32_t y0, int32_t z0, void* stream);
int sx_sw_finalize(float* preds, const float* cnt, int32_t K, int64_t V, int32_t brats, float* hard, void* stream);

/* debug knobs for bring-up (descriptor field overrides); not part of the stable ABI */
int sx_gemm_debug_set(const char* key, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* SEGTRAN_B200_H_ */
