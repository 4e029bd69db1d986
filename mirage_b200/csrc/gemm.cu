// mirage_b200/csrc/gemm.cu
//
// Host side of mb_gemm: argument checking, TMA descriptors, tile-shape / pairing heuristic and
// dispatch into the explicit kernel instantiations (gemm_{single,pair}_{a,b}.cu; kernels in
// gemm_impl.cuh).  Replaces the nn.Linear / Conv2d patch-embedding call sites of the reference; see
// include/mirage_b200.h.
#include "gemm_impl.cuh"

namespace mb200 {

int dispatch_gemm_single_a(int, int, int, const CUtensorMap&, const CUtensorMap&, const GemmDev&, cudaStream_t);
int dispatch_gemm_single_b(int, int, int, const CUtensorMap&, const CUtensorMap&, const GemmDev&, cudaStream_t);
int dispatch_gemm_pair_a(int, int, int, const CUtensorMap&, const CUtensorMap&, const GemmDev&, cudaStream_t);
int dispatch_gemm_pair_b(int, int, int, const CUtensorMap&, const CUtensorMap&, const GemmDev&, cudaStream_t);

int dispatch_gemm_single(int bn, int layout, int epi, const CUtensorMap& ta, const CUtensorMap& tb,
                         const GemmDev& p, cudaStream_t stream) {
  int rc = dispatch_gemm_single_a(bn, layout, epi, ta, tb, p, stream);
  if (rc == 1) rc = dispatch_gemm_single_b(bn, layout, epi, ta, tb, p, stream);
  return rc;
}

int dispatch_gemm_pair(int bn, int layout, int epi, const CUtensorMap& ta, const CUtensorMap& tb,
                       const GemmDev& p, cudaStream_t stream) {
  int rc = dispatch_gemm_pair_a(bn, layout, epi, ta, tb, p, stream);
  if (rc == 1) rc = dispatch_gemm_pair_b(bn, layout, epi, ta, tb, p, stream);
  return rc;
}

// Estimated time ~ waves * BN * penalty: every SM owns 128 output rows of a tile in both kernels (a
// pair tile is 256 rows over 2 SMs), so the per-SM work of one wave is proportional to BN.  Pair tiles
// halve the B traffic per SM; the 1-CTA kernel is L2-operand-bound at BN=128 and shared-memory-bound at
// BN=64, but offers twice as many, smaller tiles for problems that cannot fill 74 CTA pairs.
// row_stats[m][0] = sum over the partial slots 1..parts, in slot order (deterministic; the epilogues never use atomics)
__global__ void __launch_bounds__(256)
ln_stats_reduce_kernel(float* __restrict__ stats, int M, int parts) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  float2* sp = reinterpret_cast<float2*>(stats) + static_cast<long long>(row) * (parts + 1);
  float2 t = sp[1];
  for (int q = 2; q <= parts; ++q) {
    const float2 v = sp[q];
    t.x += v.x;
    t.y += v.y;
  }
  sp[0] = t;
}

static void pick_tiling(long long m, long long n, int splits, int force_bn, int force_pair, bool pair_ok,
                        int* bn_out, bool* pair_out) {
  const int sms = sm_count();
  double best = 1e30;
  *bn_out = 128;
  *pair_out = false;
  struct Cand { int bn; bool pair; double penalty; };
  // penalties measured on B200 at M = 25344 (scripts/check_gemm.py perfpre): a 128-wide pair tile runs at
  // 0.72-0.75 of the 256-wide one on the same problem; 1-CTA tiles at 0.65 (L2 operand traffic)
  const Cand cands[4] = {{256, true, 1.0}, {128, true, 1.36}, {128, false, 1.55}, {64, false, 2.0}};
  for (const Cand& c : cands) {
    if (force_bn && c.bn != force_bn) continue;
    if (force_pair == 1 && c.pair) continue;
    if (force_pair == 2 && !c.pair) continue;
    if (c.pair && !pair_ok) continue;
    const long long bm = c.pair ? 256 : 128;
    const long long tiles = ((m + bm - 1) / bm) * ((n + c.bn - 1) / c.bn) * splits;
    const long long slots = c.pair ? sms / 2 : sms;
    const long long waves = (tiles + slots - 1) / slots;
    const double cost = double(waves) * c.bn * c.penalty;
    if (cost < best) {
      best = cost;
      *bn_out = c.bn;
      *pair_out = c.pair;
    }
  }
}

}  // namespace mb200

using namespace mb200;

extern "C" int mb_gemm_ln_parts(int64_t m, int64_t n) {
  int bn;
  bool pair;
  pick_tiling(m, n, 1, 0, 0, true, &bn, &pair);
  return static_cast<int>((n + bn / 2 - 1) / (bn / 2));
}

extern "C" int mb_gemm(const mb_gemm_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(a != nullptr, "mb_gemm: null args");
  MB_REQUIRE(a->m > 0 && a->n > 0 && a->k > 0, "mb_gemm: empty problem m=%lld n=%lld k=%lld",
             (long long)a->m, (long long)a->n, (long long)a->k);
  MB_REQUIRE(a->m < (1ll << 31) && a->n < (1ll << 31) && a->k < (1ll << 31), "mb_gemm: too large");
  MB_REQUIRE(a->n % 8 == 0, "mb_gemm: N=%lld must be a multiple of 8", (long long)a->n);
  MB_REQUIRE(a->ldc % 8 == 0 || (a->epilogue & MB_EPI_UNPATCH),
             "mb_gemm: ldc=%lld must be a multiple of 8", (long long)a->ldc);
  MB_REQUIRE(a->a && a->b && a->out, "mb_gemm: null operand pointer");
  MB_REQUIRE(((reinterpret_cast<uintptr_t>(a->a) | reinterpret_cast<uintptr_t>(a->b) |
               reinterpret_cast<uintptr_t>(a->out) | reinterpret_cast<uintptr_t>(a->residual) |
               reinterpret_cast<uintptr_t>(a->bias) | reinterpret_cast<uintptr_t>(a->aux_in) |
               reinterpret_cast<uintptr_t>(a->aux_out)) & 15) == 0,
             "mb_gemm: operand, output, bias, residual and aux pointers must be 16-byte aligned "
             "(TMA boxes, 16-byte epilogue accesses, vector reductions)");
  MB_REQUIRE(a->k_splits >= 1, "mb_gemm: k_splits must be >= 1");
  MB_REQUIRE(a->k_splits == 1 || (a->epilogue & MB_EPI_ATOMIC),
             "mb_gemm: k_splits > 1 requires MB_EPI_ATOMIC");
  MB_REQUIRE(!(a->epilogue & MB_EPI_ATOMIC) || a->out_dtype == MB_F32,
             "mb_gemm: MB_EPI_ATOMIC requires an f32 output");
  MB_REQUIRE(!(a->epilogue & MB_EPI_DGELU) || a->aux_in, "mb_gemm: MB_EPI_DGELU needs aux_in");
  MB_REQUIRE((a->epilogue & ~(MB_EPI_GELU | MB_EPI_DGELU | MB_EPI_ATOMIC | MB_EPI_UNPATCH)) == 0,
             "mb_gemm: unknown epilogue bits 0x%x", a->epilogue);
  if (a->residual) MB_REQUIRE(a->ld_res % 4 == 0, "mb_gemm: ld_res must be a multiple of 4");
  MB_REQUIRE(a->colsum_out == nullptr ||
                 ((a->epilogue & MB_EPI_DGELU) && a->out_dtype == MB_BF16 && !a->residual &&
                  !(a->epilogue & (MB_EPI_ATOMIC | MB_EPI_UNPATCH | MB_EPI_GELU)) && a->out_row_period == 0 &&
                  (reinterpret_cast<uintptr_t>(a->colsum_out) & 15) == 0),
             "mb_gemm: colsum_out needs the plain MB_EPI_DGELU epilogue with a bf16 output and a 16-byte aligned buffer");
  if (a->aux_in || a->aux_out) MB_REQUIRE(a->ld_aux % 8 == 0, "mb_gemm: ld_aux must be a multiple of 8");
  const bool tf32 = (a->in_dtype == MB_F32);
  const int esize = tf32 ? 4 : 2;
  const int bk = 128 / esize;

  int layout;
  if (a->a_layout == MB_MAJOR_K && a->b_layout == MB_MAJOR_K) layout = tf32 ? LAY_KK_TF32 : LAY_KK_BF16;
  else if (a->a_layout == MB_MAJOR_K && a->b_layout == MB_MAJOR_MN) layout = LAY_KMN_BF16;
  else if (a->a_layout == MB_MAJOR_MN && a->b_layout == MB_MAJOR_MN) layout = LAY_MNMN_BF16;
  else if (a->a_layout == MB_A_PATCH32 && a->b_layout == MB_MAJOR_K) layout = LAY_PATCH_TF32;
  else MB_REQUIRE(false, "mb_gemm: unsupported layout combination a=%d b=%d", a->a_layout, a->b_layout);
  MB_REQUIRE(layout == LAY_KK_TF32 || layout == LAY_PATCH_TF32 || !tf32,
             "mb_gemm: MN-major operands require bf16");

  int bn;
  bool pair;
  pick_tiling(a->m, a->n, a->k_splits, a->block_n, a->cta_pair, true, &bn, &pair);
  MB_REQUIRE((pair && (bn == 256 || bn == 128)) || (!pair && (bn == 128 || bn == 64)),
             "mb_gemm: block_n=%d / cta_pair=%d is not an available tiling (pairs: 256|128, single: 128|64)",
             a->block_n, a->cta_pair);

  GemmDev p;
  p.out = a->out;
  p.bias = a->bias;
  p.residual = a->residual;
  p.aux_in = reinterpret_cast<const __nv_bfloat16*>(a->aux_in);
  p.aux_out = reinterpret_cast<__nv_bfloat16*>(a->aux_out);
  p.colsum_out = a->colsum_out;
  p.twin_out = reinterpret_cast<__nv_bfloat16*>(a->twin_out);
  p.ld_twin = a->ld_twin;
  p.row_stats = a->row_stats;
  p.ln_stats = a->ln_stats;
  p.ln_c1 = a->ln_c1;
  p.ln_eps = a->ln_eps;
  p.ln_parts = a->ln_parts;
  p.M = (int)a->m;
  p.N = (int)a->n;
  p.K = (int)a->k;
  p.ldc = a->ldc;
  p.ld_res = a->ld_res;
  p.ld_aux = a->ld_aux;
  p.res_period = (int)a->res_period;
  p.out_f32 = (a->out_dtype == MB_F32);
  p.epilogue = a->epilogue;
  const int bm = pair ? 256 : kBM;
  p.m_tiles = (p.M + bm - 1) / bm;
  p.n_tiles = (p.N + bn - 1) / bn;
  p.k_blocks = (p.K + bk - 1) / bk;  // 64 bf16 or 32 tf32 elements of K per stage, all layouts
  p.k_splits = a->k_splits > p.k_blocks ? p.k_blocks : a->k_splits;
  p.kb_per_split = (p.k_blocks + p.k_splits - 1) / p.k_splits;
  p.k_splits = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.total_tiles = p.m_tiles * p.n_tiles * p.k_splits;
  // row direction: opposite to the producer of A for the token-major GEMMs; split-K (wgrad) walks the token
  // dimension inside its K loop and keeps a fixed order
  if (a->k_splits == 1 && !(a->epilogue & MB_EPI_ATOMIC) && a->a_layout == MB_MAJOR_K) {
    p.m_reverse = take_direction() < 0 ? 1 : 0;
  } else {
    p.m_reverse = 0;
    note_direction(+1);
  }
  p.rows_per_img = 0;
  p.grid_w = 0;
  p.up_c = a->up_channels;
  p.up_ph = a->up_ph;
  p.up_pw = a->up_pw;
  p.up_gh = a->up_gh;
  p.up_gw = a->up_gw;
  if (a->epilogue & MB_EPI_UNPATCH) {
    MB_REQUIRE(a->up_channels > 0 && a->up_ph > 0 && a->up_pw > 0 && a->up_gh > 0 && a->up_gw > 0,
               "mb_gemm: MB_EPI_UNPATCH needs up_* geometry");
    MB_REQUIRE(a->up_pw % 8 == 0, "mb_gemm: MB_EPI_UNPATCH needs patch width %% 8 == 0");
    MB_REQUIRE(a->n == (int64_t)a->up_channels * a->up_ph * a->up_pw,
               "mb_gemm: MB_EPI_UNPATCH needs N = C*ph*pw");
    MB_REQUIRE(a->m % ((int64_t)a->up_gh * a->up_gw) == 0,
               "mb_gemm: MB_EPI_UNPATCH needs M multiple of gh*gw");
    MB_REQUIRE(a->out_row_period == 0, "mb_gemm: MB_EPI_UNPATCH excludes the output row map");
  }
  p.orow_period = (int)a->out_row_period;
  p.orow_stride = a->out_row_stride;
  p.orow_offset = a->out_row_offset;

  // ---- epilogue variant
  int epi = EPI_GENERIC;
  const bool special = (a->epilogue & MB_EPI_UNPATCH) || a->out_row_period > 0 ||
                       (a->residual && a->res_period > 0);
  if (special && p.out_f32 && !(a->epilogue & (MB_EPI_GELU | MB_EPI_DGELU | MB_EPI_ATOMIC)) &&
      !((a->epilogue & MB_EPI_UNPATCH) && a->up_pw % 4 != 0)) {
    epi = EPI_MAP;  // bias (+ residual / pos-emb rows) -> fp32 through a row map or the un-patchify map
  }
  if (!special) {
    if (a->epilogue & MB_EPI_GELU) {
      if (!p.out_f32 && !a->residual && !(a->epilogue & (MB_EPI_DGELU | MB_EPI_ATOMIC))) epi = EPI_GELU;
    } else if (a->epilogue & MB_EPI_DGELU) {
      if (!p.out_f32 && !a->residual && !(a->epilogue & MB_EPI_ATOMIC)) epi = EPI_DGELU;
    } else if (a->residual) {
      if (p.out_f32 && !(a->epilogue & MB_EPI_ATOMIC)) epi = EPI_RES;
    } else if (p.out_f32) {
      epi = EPI_F32;
    } else {
      epi = EPI_BF16;
    }
  }

  // folded LayerNorm (inference path): dedicated epilogues, plain K-major bf16 GEMMs only
  const bool ln_any = a->twin_out != nullptr || a->ln_stats != nullptr;
  if (ln_any) {
    MB_REQUIRE(layout == LAY_KK_BF16 && !special && a->k_splits == 1 && !(a->epilogue & (MB_EPI_DGELU | MB_EPI_ATOMIC)),
               "mb_gemm: the folded-LayerNorm epilogues need a plain K-major bf16 GEMM");
    MB_REQUIRE(!(a->twin_out && a->ln_stats), "mb_gemm: twin_out and ln_stats are mutually exclusive");
    if (a->twin_out) {
      MB_REQUIRE(a->residual && p.out_f32 && a->row_stats && !(a->epilogue & MB_EPI_GELU) && a->ld_twin % 8 == 0 &&
                     (reinterpret_cast<uintptr_t>(a->twin_out) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(a->row_stats) & 7) == 0,
                 "mb_gemm: twin_out needs the f32 residual epilogue, row_stats and a 16-byte aligned twin");
      const int parts = static_cast<int>((a->n + bn / 2 - 1) / (bn / 2));
      MB_REQUIRE(a->ln_parts == parts,
                 "mb_gemm: row_stats laid out for %d partial sums per row, this tiling (block_n %d) produces %d -- "
                 "size it with mb_gemm_ln_parts()", a->ln_parts, bn, parts);
      epi = EPI_RES_LN;
    } else {
      MB_REQUIRE(a->ln_parts >= 1 && a->ln_parts <= 64, "mb_gemm: ln_parts %d out of range", a->ln_parts);
      MB_REQUIRE(a->ln_c1 && a->bias && !p.out_f32 && !a->residual && !a->aux_out &&
                     (reinterpret_cast<uintptr_t>(a->ln_stats) & 7) == 0 && (reinterpret_cast<uintptr_t>(a->ln_c1) & 15) == 0,
                 "mb_gemm: ln_stats needs ln_c1, bias, a bf16 output and no residual / aux_out");
      epi = (a->epilogue & MB_EPI_GELU) ? EPI_GELU_LN : EPI_BF16_LN;
    }
  }

  CUtensorMap ta, tb;
  const TmaDtype tdt = tf32 ? kTmaF32 : kTmaBF16;
  // ---- A
  if (a->a_layout == MB_MAJOR_K) {
    uint64_t dims[2] = {(uint64_t)a->k, (uint64_t)a->m};
    uint64_t str[1] = {(uint64_t)a->lda * esize};
    uint32_t box[2] = {(uint32_t)bk, (uint32_t)kBM};
    if (make_tensor_map(&ta, a->a, tdt, 2, dims, str, box)) return -1;
  } else if (a->a_layout == MB_MAJOR_MN) {
    uint64_t dims[2] = {(uint64_t)a->m, (uint64_t)a->k};
    uint64_t str[1] = {(uint64_t)a->lda * esize};
    uint32_t box[2] = {64u, 64u};
    if (make_tensor_map(&ta, a->a, tdt, 2, dims, str, box)) return -1;
  } else {
    MB_REQUIRE(tf32, "mb_gemm: MB_A_PATCH32 requires in_dtype = MB_F32");
    MB_REQUIRE(a->img_h % 32 == 0 && a->img_w % 32 == 0, "mb_gemm: image %dx%d not divisible by 32",
               a->img_h, a->img_w);
    const int gh = a->img_h / 32, gw = a->img_w / 32;
    MB_REQUIRE(128 % gw == 0 && (gh * gw) % 128 == 0,
               "mb_gemm: MB_A_PATCH32 needs grid width dividing 128 and >=128 patches (got %dx%d)",
               gh, gw);
    MB_REQUIRE(a->k == 1024, "mb_gemm: MB_A_PATCH32 requires K = 1024");
    MB_REQUIRE(a->m % (gh * gw) == 0, "mb_gemm: M not a multiple of patches per image");
    const uint64_t nimg = (uint64_t)(a->m / (gh * gw));
    // innermost first: pw, nw, ph, nh, b
    uint64_t dims[5] = {32, (uint64_t)gw, 32, (uint64_t)gh, nimg};
    uint64_t str[4] = {32ull * 4, (uint64_t)a->img_w * 4, (uint64_t)a->img_w * 32 * 4,
                       (uint64_t)a->img_w * a->img_h * 4};
    uint32_t box[5] = {32u, (uint32_t)gw, 1u, (uint32_t)(128 / gw), 1u};
    if (make_tensor_map(&ta, a->a, kTmaF32, 5, dims, str, box)) return -1;
    p.rows_per_img = gh * gw;
    p.grid_w = gw;
  }
  // ---- B
  if (a->b_layout == MB_MAJOR_K) {
    uint64_t dims[2] = {(uint64_t)a->k, (uint64_t)a->n};
    uint64_t str[1] = {(uint64_t)a->ldb * esize};
    uint32_t box[2] = {(uint32_t)bk, (uint32_t)(pair ? bn / 2 : bn)};
    if (make_tensor_map(&tb, a->b, tdt, 2, dims, str, box)) return -1;
  } else {
    uint64_t dims[2] = {(uint64_t)a->n, (uint64_t)a->k};
    uint64_t str[1] = {(uint64_t)a->ldb * esize};
    uint32_t box[2] = {64u, 64u};
    if (make_tensor_map(&tb, a->b, tdt, 2, dims, str, box)) return -1;
  }

  if (ln_any) {
    const int rc_ln = dispatch_gemm_ln(bn, pair, epi, ta, tb, p, stream);
    MB_REQUIRE(rc_ln != 1, "mb_gemm: no folded-LayerNorm kernel for block_n %d, pair %d", bn, (int)pair);
    if (rc_ln == 0 && epi == EPI_RES_LN) {
      ln_stats_reduce_kernel<<<static_cast<unsigned>((a->m + 255) / 256), 256, 0, stream>>>(
          a->row_stats, static_cast<int>(a->m), a->ln_parts);
      MB_CHECK_CUDA(cudaGetLastError());
    }
    return rc_ln;
  }
  int rc = pair ? dispatch_gemm_pair(bn, layout, epi, ta, tb, p, stream)
                : dispatch_gemm_single(bn, layout, epi, ta, tb, p, stream);
  if (rc == 1 && epi != EPI_GENERIC)
    rc = pair ? dispatch_gemm_pair(bn, layout, EPI_GENERIC, ta, tb, p, stream)
              : dispatch_gemm_single(bn, layout, EPI_GENERIC, ta, tb, p, stream);
  MB_REQUIRE(rc != 1, "mb_gemm: no kernel for layout %d, block_n %d, pair %d", layout, bn, (int)pair);
  return rc;
}
