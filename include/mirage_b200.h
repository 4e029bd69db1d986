/* include/mirage_b200.h
 *
 * C ABI of libmirage_b200.so -- the B200 (sm_100a) kernels behind the MIRAGE MultiViT hot path.
 *
 * The reference (j-morano/MIRAGE) has no FFI of its own: every op on this path is a PyTorch
 * library call.  Each entry point below therefore cites the reference call site (file:line under
 * the reference checkout) whose arithmetic it replaces.  The host side (mirage_b200/*.py) binds
 * these with ctypes and exposes the reference's own nn.Module API on top.
 *
 * Conventions
 *   - every function returns 0 on success, a negative code on failure; mb_last_error() then holds
 *     a human-readable message (thread-local).
 *   - all pointers are DEVICE pointers unless stated otherwise; nothing is allocated inside.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised.
 *   - row-major tensors; "ld" arguments are leading dimensions in ELEMENTS.
 *   - bf16 = __nv_bfloat16 bit pattern (uint16_t), f32 = float, i64 = int64_t.
 */
#ifndef MIRAGE_B200_H_
#define MIRAGE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_VERSION 1

/* ---------------------------------------------------------------- runtime ------------------ */
const char* mb_last_error(void);
int mb_version(void);
int mb_sm_count(void);          /* SMs the persistent kernels size their grids to (physical - reserve) */
/* Keep `n` SMs free of persistent CTAs (GEMM, attention, LayerNorm backward size their grids to the rest), so
 * that communication kernels running concurrently (the NCCL gradient all-reduce launched from inside backward)
 * find SMs without waiting for -- or delaying -- a whole compute wave.  Returns the previous value. */
int mb_set_sm_reserve(int n);
/* Programmatic dependent launch between this library's kernels (the prologue of kernel k+1 -- barrier init, TMEM
 * allocation, descriptor prefetch -- overlaps the drain of kernel k; every such kernel executes
 * griddepcontrol.wait before touching global memory, so results are unchanged).  `mode` is a bit mask: 1 = tensor
 * kernels (GEMM, attention), 2 = row kernels (LayerNorm, column sums, casts, attention tail rows).  Initial value:
 * environment MB_PDL (0..3), else 0.  Returns the previous setting. */
int mb_set_pdl(int mode);
void mb_clear_tensor_map_cache(void);

/* ---------------------------------------------------------------- GEMM --------------------- */
/* One tcgen05/TMEM GEMM with fused epilogues:  D[M,N] = epilogue( A[M,K] * B[N,K]^T ).
 *
 * Replaces every nn.Linear / patch-embedding Conv2d on the path and their backward GEMMs:
 *   Attention.qkv / .proj        mirage/utils.py:177-186
 *   Mlp.fc1 (+GELU) / .fc2       mirage/utils.py:155-157
 *   CrossAttention.q/.kv/.proj   mirage/utils.py:209-221
 *   proj_context / out_proj      mirage/output_adapters.py:272,288
 *   PatchedInputAdapter.proj     mirage/input_adapters.py:101   (a_layout = MB_A_PATCH32)
 *   SemSegInputAdapter.proj      mirage/input_adapters.py:229
 *
 * Operand layouts ("major" = which dimension is contiguous in memory):
 *   MB_MAJOR_K   A is [M, K] with K contiguous (lda = row stride).   B is [N, K], K contiguous.
 *   MB_MAJOR_MN  A is [K, M] with M contiguous (lda = row stride).   B is [K, N], N contiguous.
 * so   forward  y = x W^T      : A=x   (K-major),  B=W  (K-major)
 *      dgrad    dx = dy W      : A=dy  (K-major),  B=W  (MN-major, reduction over W's rows)
 *      wgrad    dW = dy^T x    : A=dy  (MN-major), B=x  (MN-major), k_splits > 1 allowed.
 *   MB_A_PATCH32: A is an fp32 image batch [B, 1, H, W]; row m = (b, nh, nw) is the 32x32 patch,
 *      K = 1024 in (ph, pw) order; loaded patch-row by patch-row with a 5-D TMA box (no im2col).
 *      Requires in_dtype = MB_F32 (tf32 tensor-core math), (W/32) dividing 128.
 *
 * Epilogue, applied in this order on the fp32 accumulator acc[m, n]:
 *   1. if bias:            acc += bias[n]                                   (f32 [N])
 *   2. if MB_EPI_GELU:     (aux_out ? aux_out[m,n] = bf16(acc) : -);  acc = GELU(acc)
 *   3. if MB_EPI_DGELU:    acc *= GELU'(aux_in[m,n])                         (bf16 [M, ld_aux])
 *      GELU is nn.GELU()'s exact erf form evaluated through a tanh-form fit (max abs error 3e-5, derivative
 *      1.2e-4; csrc/common.cuh gelu_fast2) -- below the bf16 rounding of the stored result.
 *   4. if residual:        acc += residual[(res_period ? m % res_period : m), n]   (f32, ld_res)
 *   5. store: out_dtype MB_BF16 or MB_F32 at out[r(m) * ldc + n];  with MB_EPI_ATOMIC (required
 *      when k_splits > 1) the store is an fp32 atomic add into a pre-zeroed (or accumulating) buffer.
 *      r(m) = m, or with out_row_period = P > 0:  r(m) = (m / P) * out_row_stride + m % P +
 *      out_row_offset -- this lets each input adapter write its tokens straight into its slice of
 *      the concatenated [B, N_all (+ global), D] token buffer (the torch.cat of model.py:384/:520).
 *      With MB_EPI_UNPATCH the store un-patchifies instead (output_adapters.py:291-294):
 *      row m = (b, nh, nw), column n = (c, py, px) lands at image pixel
 *      out[b, c, nh*up_ph + py, nw*up_pw + px] of a [B, up_channels, up_gh*up_ph, up_gw*up_pw] tensor.
 */
enum { MB_BF16 = 0, MB_F32 = 1 };
enum { MB_MAJOR_K = 0, MB_MAJOR_MN = 1, MB_A_PATCH32 = 2 };
enum { MB_EPI_GELU = 1, MB_EPI_DGELU = 2, MB_EPI_ATOMIC = 4, MB_EPI_UNPATCH = 8 };

typedef struct mb_gemm_args {
  const void* a;      /* bf16 (or f32 when in_dtype == MB_F32) */
  const void* b;      /* same dtype as a */
  void* out;          /* bf16 or f32, see out_dtype */
  const float* bias;  /* f32 [N] or NULL */
  const float* residual; /* f32 or NULL */
  const void* aux_in; /* bf16 [M, ld_aux] pre-activation, MB_EPI_DGELU only */
  void* aux_out;      /* bf16 [M, ld_aux] pre-activation copy, MB_EPI_GELU only, may be NULL */
  int64_t m, n, k;
  int64_t lda, ldb, ldc, ld_res, ld_aux;
  int64_t res_period; /* 0 = residual indexed by m */
  int32_t a_layout;   /* MB_MAJOR_K | MB_MAJOR_MN | MB_A_PATCH32 */
  int32_t b_layout;   /* MB_MAJOR_K | MB_MAJOR_MN */
  int32_t in_dtype;   /* MB_BF16 | MB_F32 (tf32 math) */
  int32_t out_dtype;  /* MB_BF16 | MB_F32 */
  int32_t epilogue;   /* bitmask of MB_EPI_* */
  int32_t k_splits;   /* >= 1 */
  int32_t block_n;    /* 0 = auto, else 64 / 128 / 256 */
  int32_t img_h, img_w; /* MB_A_PATCH32 only */
  int64_t out_row_period, out_row_stride, out_row_offset; /* 0,0,0 = identity row map */
  int32_t up_channels, up_ph, up_pw, up_gh, up_gw;        /* MB_EPI_UNPATCH only */
  int32_t cta_pair; /* 0 = auto, 1 = one CTA per 128-row tile, 2 = CTA pairs (cta_group::2, 256 rows) */
  float* colsum_out; /* f32 [N] or NULL: ACCUMULATES the column sums of the stored output (the bias gradient
                        of the layer whose data gradient this GEMM produces); MB_EPI_DGELU with bf16 out only */
  /* LayerNorm folded into the GEMMs around it (inference path; Block.norm1 / norm2, mirage/utils.py:260-261):
   *   producer  (residual GEMM, f32 out):  twin_out receives bf16(out) and row_stats[m][1 + part] = {sum_n out,
   *             sum_n out^2} over the part's columns, then row_stats[m][0] = their sum in slot order (f32
   *             [M, 1 + ln_parts, 2], every slot written, no atomics: deterministic) -- what the following LayerNorm
   *             needs, without re-reading out;
   *   consumer  (A = that bf16 twin, B = bf16(gamma * W)):  out = rstd[m] * acc - rstd[m] * mean[m] * ln_c1[n] + bias[n]
   *             with mean / rstd from ln_stats[m][0] (f32 [M, 1 + ln_parts, 2] = the producer's row_stats, statistics over K = the
   *             LayerNorm width), ln_c1[n] = sum_k bf16(gamma_k W_nk), bias[n] = b[n] + sum_k beta_k W_nk;
   *             then GELU with MB_EPI_GELU.  Equals LayerNorm(x) W^T + b up to bf16 rounding. */
  void* twin_out;        /* bf16 [M, ld_twin] or NULL */
  int64_t ld_twin;
  float* row_stats;      /* f32 [M, 1 + ln_parts, 2]; required with twin_out */
  const float* ln_stats; /* f32 [M, 1 + ln_parts, 2] or NULL */
  const float* ln_c1;    /* f32 [N]; required with ln_stats */
  float ln_eps;
  int32_t ln_parts;      /* partial sums per row in row_stats / ln_stats: mb_gemm_ln_parts(m, n) of the PRODUCER GEMM */
} mb_gemm_args;

int mb_gemm(const mb_gemm_args* args, void* stream);
/* Number of partial sums per row the folded-LayerNorm producer epilogue writes for an [m, n] output with the tiling
 * mb_gemm picks for it (one per block_n / 2 columns); row_stats / ln_stats are f32 [m, 1 + parts, 2]. */
int mb_gemm_ln_parts(int64_t m, int64_t n);

/* ---------------------------------------------------------------- attention ---------------- */
/* Fused softmax(Q K^T * scale) V, no mask, no dropout (the reference always runs attn_drop = 0).
 * Replaces F.scaled_dot_product_attention AND the following .transpose(1, 2).reshape(B, N, C) copy:
 *   Attention.forward        mirage/utils.py:181-185
 *   CrossAttention.forward   mirage/utils.py:216-220
 *
 * q  : bf16, token-major: element (b, i, h, d) at q[(b*nq + i)*ldq + h*head_dim + d]
 * k,v: bf16, same convention with nk / ldk / ldv.  (For the fused qkv Linear output [B*N, 3*D] pass
 *      q = base, k = base + D, v = base + 2*D and ldq = ldk = ldv = 3*D.)
 * out: bf16 [B*nq, ldo] with head h in columns [h*head_dim, (h+1)*head_dim) -- i.e. already in the
 *      layout the output projection consumes.
 * lse: optional f32 [B, heads, nq]: log-sum-exp of the scaled scores (saved for the backward pass).
 * head_dim is 64 (encoder) or 32 (decoder).
 */
typedef struct mb_attn_args {
  const void* q;
  const void* k;
  const void* v;
  void* out;
  float* lse;
  int64_t batch, heads, nq, nk;
  int64_t ldq, ldk, ldv, ldo;
  int32_t head_dim;
  float scale;
} mb_attn_args;

int mb_attn_fwd(const mb_attn_args* args, void* stream);

/* Attention backward (autograd of the call sites above).  Inputs: the forward's q/k/v/out/lse and
 * d_out (bf16, same layout as out).  Outputs dq/dk/dv: bf16, token-major like q/k/v (they may be
 * column slices of one fused [B*N, 3*D] gradient buffer).  When nk > 128 a workspace of
 * mb_attn_bwd_workspace() bytes is required (fp32 dQ accumulation across key blocks). */
typedef struct mb_attn_bwd_args {
  const void* q;
  const void* k;
  const void* v;
  const void* out;
  const void* d_out;
  const float* lse;
  void* dq;
  void* dk;
  void* dv;
  void* workspace;
  int64_t batch, heads, nq, nk;
  int64_t ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  int32_t head_dim;
  float scale;
} mb_attn_bwd_args;

int64_t mb_attn_bwd_workspace(int64_t batch, int64_t heads, int64_t nq, int64_t nk, int32_t head_dim);
int mb_attn_bwd(const mb_attn_bwd_args* args, void* stream);

/* ---------------------------------------------------------------- row kernels (HBM-bound) -- */
/* LayerNorm(eps) with affine, fp32 statistics.  nn.LayerNorm at mirage/utils.py:241,250,260-261,
 * output_adapters.py:115-117, mirage_wrapper.py:201.
 * x f32 [rows, ldx]; y bf16 or f32 [rows, ldy] (y_dtype); mean/rstd optional f32 [rows] (saved for
 * the backward pass).  dim must be a multiple of 128, at most 1024. */
int mb_layernorm_fwd(const float* x, const float* weight, const float* bias, void* y,
                     int32_t y_dtype, float* mean, float* rstd, int64_t rows, int64_t dim,
                     int64_t ldx, int64_t ldy, float eps, void* stream);

/* LayerNorm backward.  dy bf16 or f32 (dy_dtype); dx f32 = LN'(dy) (+ dres if non-NULL: the gradient
 * arriving through the residual branch); dx_bf16 (optional, contiguous bf16 [rows, dim]) receives the
 * same values rounded to bf16 -- the operand of the weight/data-gradient GEMMs that follow;
 * dweight/dbias f32 [dim], overwritten or accumulated.
 * dx_colsum (optional, f32 [dim], overwritten or accumulated per dx_colsum_accumulate): column sums of dx = the
 * bias gradient of the nn.Linear whose output gradient dx is (Attention.proj / Mlp.fc2, utils.py:157,186) --
 * saves a separate mb_colsum pass over dx.
 * workspace: mb_layernorm_bwd_workspace(rows, dim) bytes of device memory. */
int64_t mb_layernorm_bwd_workspace(int64_t rows, int64_t dim);
int mb_layernorm_bwd(const void* dy, int32_t dy_dtype, const float* x, const float* weight,
                     const float* mean, const float* rstd, const float* dres, float* dx,
                     void* dx_bf16, float* dweight, float* dbias, int32_t accumulate,
                     void* workspace, int64_t rows, int64_t dim, int64_t ldx, int64_t lddy,
                     int64_t lddx, float* dx_colsum, int32_t dx_colsum_accumulate, void* stream);

/* out[c] (+)= sum_r a[r, c] -- bias gradient of every nn.Linear on the path. */
int64_t mb_colsum_workspace(int64_t rows, int64_t cols);
int mb_colsum(const void* a, int32_t a_dtype, float* out, int32_t accumulate, void* workspace,
              int64_t rows, int64_t cols, int64_t lda, void* stream);

/* Visible-token selection, mirage/model.py:384-391:
 *   out[b, j, :]          = src[b, ids_keep[b, j], :]   j < n_keep      (bit-exact row copy)
 *   out[b, n_keep + t, :] = global_tokens[t, :]                          (global tokens LAST)
 * src f32 [B, n_src, dim], ids_keep i64 [B, n_keep], out f32 [B, n_keep + n_global, dim]. */
int mb_token_gather_fwd(const float* src, const int64_t* ids_keep, const float* global_tokens,
                        float* out, int64_t batch, int64_t n_src, int64_t n_keep, int64_t n_global,
                        int64_t dim, void* stream);
/* Backward: dsrc (zero-filled here) receives dout rows at ids_keep; dglobal[t] = sum_b dout[b, n_keep+t]. */
int mb_token_gather_bwd(const float* dout, const int64_t* ids_keep, float* dsrc, float* dglobal,
                        int64_t batch, int64_t n_src, int64_t n_keep, int64_t n_global, int64_t dim,
                        void* stream);
/* out[b, row_offset + t, :] = global_tokens[t, :] for the un-masked path (model.py:523-524). */
int mb_fill_global_rows(const float* global_tokens, float* out, int64_t batch, int64_t rows_total,
                        int64_t row_offset, int64_t n_global, int64_t dim, void* stream);
int mb_cast_f32_to_bf16(const float* in, void* out, int64_t n, void* stream);

/* LayerNorm + token mean-pool of the classification wrappers:
 *   pooled[b, c] = mean over t in [row_begin, row_end) of LayerNorm(x[b, t, :])[c]
 * mirage_wrapper.py:217-224 (MIRAGEClsGlobal.forward / .pool), :230-233 (CLS), :236-244 (TokenMix: two
 * calls writing the two halves of a [B, 2*dim] output through ld_pooled).
 * x f32 [B, n_tokens, dim] contiguous; pooled f32 [B, ld_pooled]; xhat_mean f32 [B, dim], mean / rstd
 * f32 [B, n_tokens] (entries of the pooled rows are written) are saved for the backward pass.
 * workspace: mb_ln_meanpool_workspace(batch, dim) bytes.  dim % 128 == 0, dim <= 1024. */
int64_t mb_ln_meanpool_workspace(int64_t batch, int64_t dim);
int mb_ln_meanpool_fwd(const float* x, const float* gamma, const float* beta, float* pooled, int64_t ld_pooled,
                       float* xhat_mean, float* mean, float* rstd, void* workspace, int64_t batch,
                       int64_t n_tokens, int64_t dim, int64_t row_begin, int64_t row_end, float eps,
                       void* stream);
/* Backward: dx f32 [B, n_tokens, dim] receives the gradient of the pooled rows (rows outside the range are
 * zeroed when zero_outside != 0, left untouched otherwise -- the second call of TokenMix); d_gamma / d_beta
 * f32 [dim] overwritten or accumulated. */
int mb_ln_meanpool_bwd(const float* d_pooled, int64_t ld_dp, const float* x, const float* gamma,
                       const float* mean, const float* rstd, const float* xhat_mean, float* dx,
                       float* d_gamma, float* d_beta, int32_t accumulate, int32_t zero_outside,
                       int64_t batch, int64_t n_tokens, int64_t dim, int64_t row_begin, int64_t row_end,
                       void* stream);

/* ---------------------------------------------------------------- optimizer step (8(f1)) ---- */
/* Fused AdamW over parameter groups + global gradient norm / clip / skip + in-place refresh of the bf16
 * weight shadows the GEMMs read.  Replaces, per training step:
 *   optim.AdamW.step over get_parameter_groups()        mutils/optim_factory.py:33-92, :171-172
 *   per-step lr / weight-decay assignment                run_pretraining.py:683-688  (hyper table)
 *   unscale_/clip_grad_norm_/get_grad_norm_/skip_grad    mutils/native_scaler.py:16-37, :46-61
 * Arithmetic is torch.optim.AdamW's (decoupled decay, bias-corrected, fp32 state).
 *
 * A SEGMENT is one parameter tensor; a block of mb_adamw_step owns 4096 consecutive elements of one
 * segment: block_prefix[i] = first block of segment i (i32 [n_segments + 1], device), n_blocks =
 * block_prefix[n_segments] = sum of mb_optim_blocks(numel_i).  hyper[group] carries the values the host
 * schedule assigns each step; state holds the step counter and the scalars shared by all blocks.
 * Call order per step:
 *   clip_norm > 0 or skip_norm > 0:  mb_sumsq (per gradient bucket, into consecutive partials) ->
 *       mb_optim_prepare(partials) -> mb_adamw_step(partials = NULL)
 *   otherwise:  mb_optim_prepare(n_partials = 0) -> mb_adamw_step(partials [n_blocks]) -> mb_optim_finish
 * Either way state->grad_norm ends up holding the L2 norm of the (unclipped) gradient. */
typedef struct mb_optim_segment {
  float* param;     /* f32 [numel] */
  float* grad;      /* f32 [numel] */
  float* exp_avg;   /* f32 [numel] */
  float* exp_avg_sq;/* f32 [numel] */
  void* shadow;     /* bf16 [numel] or NULL: rewritten with the updated parameter */
  int64_t numel;
  int32_t group;    /* index into the hyper table */
  int32_t flags;    /* bit 0: all pointers 16-byte aligned (shadow 8) -> vector path */
} mb_optim_segment;

typedef struct mb_optim_hyper {
  float lr;           /* param_group['lr'] before lr_scale */
  float lr_scale;     /* layer-wise lr decay factor (optim_factory.py:70-74) */
  float weight_decay; /* param_group['weight_decay'] */
  float pad_;
} mb_optim_hyper;

typedef struct mb_optim_state {
  int64_t step;          /* completed optimizer steps */
  float bias_corr1;      /* 1 - beta1^step */
  float bias_corr2_sqrt; /* sqrt(1 - beta2^step) */
  float clip_coef;       /* factor applied to the gradient this step (1 when not clipping) */
  float grad_norm;       /* L2 norm of the gradient before clipping */
  int32_t skipped;       /* 1 when the step was dropped (norm >= skip_norm) */
  int32_t pad_;
} mb_optim_state;

int64_t mb_optim_blocks(int64_t numel);
int mb_sumsq(const float* x, int64_t n, float* partials, int32_t n_partials, void* stream);
int mb_optim_prepare(mb_optim_state* state, const float* partials, int32_t n_partials, float clip_norm,
                     float skip_norm, float beta1, float beta2, void* stream);
int mb_adamw_step(const mb_optim_segment* segments, const int32_t* block_prefix, int32_t n_segments,
                  int64_t n_blocks, const mb_optim_hyper* hyper, const mb_optim_state* state, float beta1,
                  float beta2, float eps, int32_t zero_grad, float* partials, void* stream);
int mb_optim_finish(mb_optim_state* state, const float* partials, int64_t n_partials, void* stream);

/* ---------------------------------------------------------------- mask sampling (8(f2)) ----- */
/* MIRAGEModel.generate_random_masks (+ sample_alphas) in one kernel, one CTA per sample
 * (mirage/model.py:168-239, :145-166): Dirichlet(alphas) split of n_encoded over the tasks, random subset
 * per task, global order with the visible tokens first.  Same distribution as the reference, its own
 * documented Philox4x32-10 stream (csrc/masks.cu) keyed by `seed` and the device-resident *draw_counter,
 * which the kernel advances by one per call (so a CUDA-graph replay draws fresh masks).
 * task_counts i32 [n_tasks] and alphas f32 [n_tasks] are HOST arrays (n_tasks <= 8).
 * Outputs: task_masks i64 [B, n_all] (tasks concatenated in order; 0 = visible, 1 = masked),
 * ids_keep i64 [B, n_encoded], ids_restore i64 [B, n_all].  done_counter: u32 device scratch, zero-initialised. */
int mb_sample_masks(uint64_t seed, uint64_t* draw_counter, uint32_t* done_counter, const int32_t* task_counts,
                    const float* alphas, int32_t n_tasks, int64_t batch, int64_t n_encoded,
                    int32_t uniform_tasks, int64_t* task_masks, int64_t* ids_keep, int64_t* ids_restore,
                    void* stream);

/* ---------------------------------------------------------------- input pipeline (8(f3)) ---- */
/* DataAugmentationForMIRAGE on the device (mutils/datasets_pretrain.py:18-83, loading :172-185): raw uint8
 * [B, H, W] in, network inputs out.  params f32 [B, 8] = {flip, shift, m00, m01, m02, m10, m11, m12} per sample
 * (inverse affine matrix in torchvision's centred convention, see csrc/augment.cu).
 *   mb_augment_image : out f32 [B, 1, H, W] = affine(clip(flip(src) / 255 + shift, 0, 1)), bilinear, zero fill
 *   mb_augment_labels: out i64 [B, OH, OW]  = nearest-resize(round(affine(flip(src)))) (bilinear on the class ids,
 *                      as torchvision does for integer tensors) */
int mb_augment_image(const uint8_t* src, const float* params, float* out, int64_t batch, int32_t height,
                     int32_t width, void* stream);
int mb_augment_labels(const uint8_t* src, const float* params, int64_t* out, int64_t batch, int32_t height,
                      int32_t width, int32_t out_height, int32_t out_width, void* stream);

/* ---------------------------------------------------------------- adapters ----------------- */
/* SemSegInputAdapter (mirage/input_adapters.py:226-229): class-embedding lookup fused with patch
 * extraction.  out[m, c*P*Q + ph*Q + pw] = class_emb[labels[b, nh*P+ph, nw*Q+pw], c] (bf16), the A
 * operand of the patch-projection GEMM, m = (b, nh, nw), K order (c, ph, pw) as Conv2d's weight. */
int mb_semseg_patches(const int64_t* labels, const void* class_emb_bf16, void* out, int64_t batch,
                      int64_t height, int64_t width, int32_t patch_h, int32_t patch_w,
                      int32_t n_classes, int32_t emb_dim, void* stream);
/* d_class_emb[cls, c] = sum of d_patches over every pixel whose label is cls (f32 [n_classes, emb_dim]). */
int mb_class_emb_grad(const int64_t* labels, const void* d_patches_bf16, float* d_class_emb,
                      int64_t batch, int64_t height, int64_t width, int32_t patch_h, int32_t patch_w,
                      int32_t n_classes, int32_t emb_dim, void* stream);

/* Row-list forms of the two calls above (visible-token embedding, below): output / gradient row r belongs to
 * source patch row_src[r] (= b * tokens_per_sample + token); a negative entry gives a zero row / is skipped. */
int mb_semseg_patches_rows(const int64_t* labels, const void* class_emb_bf16, void* out, const int32_t* row_src,
                           int64_t n_rows, int64_t batch, int64_t height, int64_t width, int32_t patch_h,
                           int32_t patch_w, int32_t n_classes, int32_t emb_dim, void* stream);
int mb_class_emb_grad_rows(const int64_t* labels, const void* d_patches_bf16, float* d_class_emb,
                           const int32_t* row_src, int64_t n_rows, int64_t batch, int64_t height, int64_t width,
                           int32_t patch_h, int32_t patch_w, int32_t n_classes, int32_t emb_dim, void* stream);

/* Visible-token embedding of MIRAGEModel.forward with masking (mirage/model.py:352-356 input adapters + :384-391
 * gather and global tokens): the masks are drawn from the token counts alone, so only the kept patches are
 * projected.  Rows t = b * (n_keep + n_global) + j, the layout the encoder consumes.
 *   mb_visible_rows      row_src [n_modalities, T] int32: source patch of row t for modality m, else -1;
 *                        row_cls [T] int32: modality of row t, n_modalities + g for global token g, -1 for an id
 *                        outside every modality.  starts / counts: HOST int32 [n_modalities] (first token of the
 *                        modality in the concatenated sequence, tokens per sample).
 *   mb_embed_rows_init   out[t] = bias_m + pos_m[token] | global_tokens[g] | 0  (f32 [T, dim]); bias / pos: HOST
 *                        arrays of n_modalities DEVICE pointers (f32 [dim], f32 [count_m, dim] or NULL).
 *   mb_gather_patches32  A[t] = the 32x32 patch of row_src[t] flattened (ph, pw) -- PatchedInputAdapter's Conv2d K
 *                        order -- as f32 and/or bf16 [T, 1024] (either may be NULL); zero rows for negative entries.
 *   mb_class_colsum      out[c] = sum of the rows of dy (f32 [T, dim]) whose row_cls is c (f32 [n_classes, dim]):
 *                        bias gradients of the adapters and the global-token gradient in one pass. */
int mb_visible_rows(const int64_t* ids_keep, int64_t batch, int64_t n_keep, int64_t n_global, int32_t n_modalities,
                    const int32_t* starts, const int32_t* counts, int32_t* row_src, int32_t* row_cls, void* stream);
int mb_embed_rows_init(const int32_t* row_src, const int32_t* row_cls, int32_t n_modalities, const int32_t* counts,
                       const float* const* bias, const float* const* pos, const float* global_tokens, float* out,
                       int64_t rows, int64_t dim, void* stream);
int mb_gather_patches32(const float* images, const int32_t* row_src, float* a_f32, void* a_bf16, int64_t rows,
                        int64_t height, int64_t width, void* stream);
int64_t mb_class_colsum_workspace(int64_t rows, int64_t dim, int32_t n_classes);
int mb_class_colsum(const float* dy, const int32_t* row_cls, float* out, void* workspace, int64_t rows, int64_t dim,
                    int32_t n_classes, void* stream);

/* SpatialOutputAdapter.get_queries_and_context (mirage/output_adapters.py:188-246), use_task_queries:
 *   queries[b, t] = (r < n_vis ? ctx[b, r] : mask_token) + emb[q_start + t],  r = ids_restore[b, q_start + t]
 *   context[b, j] = (r2 < n_vis ? ctx[b, r2] : mask_token) + emb[ids_keep[b, j]],  r2 = ids_restore[b, ids_keep[b, j]]
 *   context[b, n_vis + g] = ctx[b, n_vis + g]            (global tokens get no embedding, :240-242)
 * ctx f32 [B, n_vis + n_global, dim] (proj_context output), emb f32 [n_all, dim] = task_emb + pos_emb
 * (added in that order, :180).  Bit-exact with the reference's cat/gather/add sequence in fp32. */
int mb_dec_assemble_fwd(const float* ctx, const float* mask_token, const float* emb,
                        const int64_t* ids_keep, const int64_t* ids_restore, float* queries,
                        float* context, int64_t batch, int64_t n_vis, int64_t n_global, int64_t n_all,
                        int64_t q_start, int64_t n_q, int64_t dim, void* stream);
/* Backward; ids_restore must be the inverse permutation of the shuffle behind ids_keep (it always is
 * on the path: model.py:225-227, :380-382).  workspace: n_q * dim * 4 bytes. */
int mb_dec_assemble_bwd(const float* d_queries, const float* d_context, const int64_t* ids_keep,
                        const int64_t* ids_restore, float* d_ctx, float* d_emb, float* d_mask_token,
                        void* workspace, int64_t batch, int64_t n_vis, int64_t n_global,
                        int64_t n_all, int64_t q_start, int64_t n_q, int64_t dim, void* stream);

/* image [B, C, gh*ph, gw*pw] f32 -> tokens [B*gh*gw, C*ph*pw] bf16: backward of the un-patchify
 * (output_adapters.py:291-294), producing the dgrad/wgrad operand of out_proj. */
int mb_patchify_cast(const float* img, void* out_bf16, int64_t batch, int32_t channels, int32_t ph,
                     int32_t pw, int32_t gh, int32_t gw, void* stream);

/* ---------------------------------------------------------------- masked criteria ---------- */
/* MaskedMSELoss (mirage/criterion.py:87-117, norm_pix=False) and MaskedCrossEntropyLoss (:31-51).
 * pred/target f32 [B,C,H,W] (CE: logits f32 [B,C,H,W], target i64 [B,H,W]); mask i64 [B, (H/scale)*(W/scale)]
 * with 1 = masked-out token (contributes to the loss), or NULL (plain mean).  loss: f32 scalar;
 * coef: f32 [B] saved for the backward.  An all-zero mask yields 0.0 (the reference returns an int64
 * tensor(0) after a host sync; here nothing is read back).  workspace: mb_masked_loss_workspace() bytes. */
int64_t mb_masked_loss_workspace(int64_t batch, int64_t height, int64_t width);
int mb_masked_mse_fwd(const float* pred, const float* target, const int64_t* mask, float* loss,
                      float* coef, void* workspace, int64_t batch, int64_t channels, int64_t height,
                      int64_t width, int32_t scale, void* stream);
int mb_masked_mse_bwd(const float* pred, const float* target, const int64_t* mask, const float* coef,
                      const float* grad_out, float* dpred, int64_t batch, int64_t channels,
                      int64_t height, int64_t width, int32_t scale, void* stream);
int mb_masked_ce_fwd(const float* logits, const int64_t* target, const int64_t* mask, float* loss,
                     float* coef, void* workspace, int64_t batch, int64_t channels, int64_t height,
                     int64_t width, int32_t scale, float label_smoothing, void* stream);
int mb_masked_ce_bwd(const float* logits, const int64_t* target, const int64_t* mask,
                     const float* coef, const float* grad_out, float* dlogits, int64_t batch,
                     int64_t channels, int64_t height, int64_t width, int32_t scale,
                     float label_smoothing, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* MIRAGE_B200_H_ */
