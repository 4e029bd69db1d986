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
int mb_sm_count(void);
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
} mb_gemm_args;

int mb_gemm(const mb_gemm_args* args, void* stream);

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
 * workspace: mb_layernorm_bwd_workspace(rows, dim) bytes of device memory. */
int64_t mb_layernorm_bwd_workspace(int64_t rows, int64_t dim);
int mb_layernorm_bwd(const void* dy, int32_t dy_dtype, const float* x, const float* weight,
                     const float* mean, const float* rstd, const float* dres, float* dx,
                     void* dx_bf16, float* dweight, float* dbias, int32_t accumulate,
                     void* workspace, int64_t rows, int64_t dim, int64_t ldx, int64_t lddy,
                     int64_t lddx, void* stream);

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
