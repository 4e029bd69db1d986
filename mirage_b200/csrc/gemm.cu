// mirage_b200/csrc/gemm.cu
//
// Persistent, warp-specialised tcgen05 GEMM for sm_100a:   D[M,N] = epilogue(A * B^T).
//
//   warp 0      TMA producer   (one lane): global -> 128B-swizzled smem ring, mbarrier tx-count
//   warp 1      MMA issuer     (one lane): tcgen05.mma.cta_group::1, M=128, N=BN, K=32 bytes/instr
//   warp 2      TMEM allocator (512 / 256 / 128 columns = 2 accumulator stages of BN columns)
//   warp 3      idle
//   warps 4-11  epilogue: tcgen05.ld (32 lanes x 32 columns) -> bias / GELU / GELU' / residual ->
//               16-byte global stores (or fp32 atomics for split-K wgrad)
//
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the main loop of
// tile i+1.  Tiles are walked n-fastest so that the CTAs resident at any moment share a handful of
// A row-panels (L2 reuse) while the whole weight matrix stays L2 resident.
//
// Replaces nn.Linear / Conv2d patch embedding call sites of the reference; see include/mirage_b200.h.
#include "../../include/mirage_b200.h"
#include "common.cuh"

namespace mb200 {

struct GemmDev {
  void* out;
  const float* bias;
  const float* residual;
  const __nv_bfloat16* aux_in;
  __nv_bfloat16* aux_out;
  int M, N, K;
  long long ldc, ld_res, ld_aux;
  int res_period;
  int out_f32;
  int epilogue;
  int m_tiles, n_tiles, k_blocks, k_splits, kb_per_split;
  int total_tiles;
  // MB_A_PATCH32
  int rows_per_img;  // tokens per image
  int grid_w;        // patches per image row
  // output row map
  int orow_period;
  long long orow_stride, orow_offset;
  // MB_EPI_UNPATCH: token-major [B*gh*gw, C*ph*pw] -> image [B, C, gh*ph, gw*pw]
  int up_c, up_ph, up_pw, up_gh, up_gw;
};

constexpr int kGemmThreads = 384;
constexpr int kBM = 128;

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // +1024: alignment
};

template <int BN, int A_LAYOUT, int B_MN, int ESIZE>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const GemmDev p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int BK = 128 / ESIZE;  // elements of K per stage for K-major operands
  constexpr bool A_MN = (A_LAYOUT == MB_MAJOR_MN);
  static_assert(!(A_MN || B_MN) || ESIZE == 2, "MN-major operands are bf16 only");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 8);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_mn = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int split = tile / tiles_mn;
      const int mn = tile - split * tiles_mn;
      const int m_blk = mn / p.n_tiles;
      const int n_blk = mn - m_blk * p.n_tiles;
      const int m0 = m_blk * kBM;
      const int n0 = n_blk * BN;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.k_blocks);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        if constexpr (A_LAYOUT == MB_MAJOR_K) {
          tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BK, m0);
        } else if constexpr (A_LAYOUT == MB_MAJOR_MN) {
#pragma unroll
          for (int s = 0; s < kBM / 64; ++s)
            tma_load_2d(sa + s * 8192, &tma_a, &full_bar[stage], m0 + s * 64, kb * 64);
        } else {
          // one 32-pixel patch row (ph = kb) of 128 consecutive patches of one image
          const int img = m0 / p.rows_per_img;
          const int nh0 = (m0 - img * p.rows_per_img) / p.grid_w;
          tma_load_5d(sa, &tma_a, &full_bar[stage], 0, 0, kb, nh0, img);
        }
        if constexpr (B_MN == 0) {
          tma_load_2d(sb, &tma_b, &full_bar[stage], kb * BK, n0);
        } else {
#pragma unroll
          for (int s = 0; s < BN / 64; ++s)
            tma_load_2d(sb + s * 8192, &tma_b, &full_bar[stage], n0 + s * 64, kb * 64);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc =
        make_idesc(kBM, BN, ESIZE == 2 ? kFmtBF16 : kFmtTF32, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
    constexpr uint32_t a_lbo = A_MN ? 8192u : 0u;
    constexpr uint32_t b_lbo = B_MN ? 8192u : 0u;
    constexpr uint32_t a_kstep = A_MN ? 2048u : 32u;
    constexpr uint32_t b_kstep = B_MN ? 2048u : 32u;
    int stage = 0;
    uint32_t phase = 0;
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int split = tile / tiles_mn;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.k_blocks);
      mbar_wait(&tmem_empty_bar[acc_stage], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc_stage * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = make_smem_desc(a_addr + k * a_kstep, a_lbo, 1024);
          const uint64_t db = make_smem_desc(b_addr + k * b_kstep, b_lbo, 1024);
          const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
          if constexpr (ESIZE == 2)
            umma_f16_ss(tmem_d, da, db, idesc, acc);
          else
            umma_tf32_ss(tmem_d, da, db, idesc, acc);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(&tmem_full_bar[acc_stage]);  // accumulator complete -> epilogue
      acc_stage ^= 1;
      if (acc_stage == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;  // which half of the BN columns
    constexpr int HALF_COLS = BN / 2;
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int split = tile / tiles_mn;
      const int mn = tile - split * tiles_mn;
      const int m_blk = mn / p.n_tiles;
      const int n_blk = mn - m_blk * p.n_tiles;
      const int row = m_blk * kBM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const int n_base = n_blk * BN + half * HALF_COLS;
      const bool add_bias = (p.bias != nullptr) && (split == 0);
      const long long res_row =
          p.residual ? (long long)(p.res_period > 0 ? row % p.res_period : row) : 0;
      const long long orow =
          p.orow_period > 0
              ? (long long)(row / p.orow_period) * p.orow_stride + row % p.orow_period + p.orow_offset
              : (long long)row;

      long long up_base = 0;  // offset of pixel (b, c=0, nh*ph, nw*pw)
      if (p.epilogue & MB_EPI_UNPATCH) {
        const int per_img = p.up_gh * p.up_gw;
        const int bi = row / per_img, t = row - bi * per_img;
        const int nh = t / p.up_gw, nw = t - nh * p.up_gw;
        const long long W = (long long)p.up_gw * p.up_pw, H = (long long)p.up_gh * p.up_ph;
        up_base = ((long long)bi * p.up_c * H + (long long)nh * p.up_ph) * W + (long long)nw * p.up_pw;
      }
      mbar_wait(&tmem_full_bar[acc_stage], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < HALF_COLS / 32; ++c) {
        const int col0 = n_base + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(acc_stage * BN + half * HALF_COLS + c * 32);
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = col0 + g * 8;
          if (col >= p.N) break;  // N % 8 == 0 is required by the host wrapper
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[g * 8 + i]);
          if (add_bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
          }
          if (row_ok) {
            if (p.epilogue & MB_EPI_GELU) {
              if (p.aux_out) {
                uint4 pk;
                pk.x = pack_bf16x2(f[0], f[1]);
                pk.y = pack_bf16x2(f[2], f[3]);
                pk.z = pack_bf16x2(f[4], f[5]);
                pk.w = pack_bf16x2(f[6], f[7]);
                *reinterpret_cast<uint4*>(p.aux_out + (long long)row * p.ld_aux + col) = pk;
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = gelu_erf(f[i]);
            }
            if (p.epilogue & MB_EPI_DGELU) {
              const uint4 pk =
                  __ldg(reinterpret_cast<const uint4*>(p.aux_in + (long long)row * p.ld_aux + col));
              const float2 h0 = unpack_bf16x2(pk.x), h1 = unpack_bf16x2(pk.y);
              const float2 h2 = unpack_bf16x2(pk.z), h3 = unpack_bf16x2(pk.w);
              f[0] *= gelu_erf_grad(h0.x); f[1] *= gelu_erf_grad(h0.y);
              f[2] *= gelu_erf_grad(h1.x); f[3] *= gelu_erf_grad(h1.y);
              f[4] *= gelu_erf_grad(h2.x); f[5] *= gelu_erf_grad(h2.y);
              f[6] *= gelu_erf_grad(h3.x); f[7] *= gelu_erf_grad(h3.y);
            }
            if (p.residual && split == 0) {
              const float* r = p.residual + res_row * p.ld_res + col;
              const float4 r0 = *reinterpret_cast<const float4*>(r);
              const float4 r1 = *reinterpret_cast<const float4*>(r + 4);
              f[0] += r0.x; f[1] += r0.y; f[2] += r0.z; f[3] += r0.w;
              f[4] += r1.x; f[5] += r1.y; f[6] += r1.z; f[7] += r1.w;
            }
            long long oidx = orow * p.ldc + col;
            if (p.epilogue & MB_EPI_UNPATCH) {
              const int pp = p.up_ph * p.up_pw;
              const int ch = col / pp, rr = col - ch * pp;
              const int py = rr / p.up_pw, px = rr - py * p.up_pw;
              const long long W = (long long)p.up_gw * p.up_pw, H = (long long)p.up_gh * p.up_ph;
              oidx = up_base + ((long long)ch * H + py) * W + px;
            }
            if (p.epilogue & MB_EPI_ATOMIC) {
              float* o = reinterpret_cast<float*>(p.out) + oidx;
#pragma unroll
              for (int i = 0; i < 8; ++i) atomicAdd(o + i, f[i]);
            } else if (p.out_f32) {
              float* o = reinterpret_cast<float*>(p.out) + oidx;
              *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
              *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
            } else {
              uint4 pk;
              pk.x = pack_bf16x2(f[0], f[1]);
              pk.y = pack_bf16x2(f[2], f[3]);
              pk.z = pack_bf16x2(f[4], f[5]);
              pk.w = pack_bf16x2(f[6], f[7]);
              __nv_bfloat16* o =
                  reinterpret_cast<__nv_bfloat16*>(p.out) + oidx;
              *reinterpret_cast<uint4*>(o) = pk;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc_stage]);
      acc_stage ^= 1;
      if (acc_stage == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int BN, int A_LAYOUT, int B_MN, int ESIZE>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_kernel<BN, A_LAYOUT, B_MN, ESIZE>;
  static bool configured = false;
  if (!configured) {
    MB_CHECK_CUDA(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int pick_block_n(long long m, long long n) {
  // Estimated time ~ waves * BN * penalty; narrower tiles re-read A from smem more often per MMA
  // (BN=64 is shared-memory-bandwidth bound), wider tiles waste columns when N is ragged.
  const int sms = sm_count();
  const long long m_tiles = (m + kBM - 1) / kBM;
  int best = 256;
  double best_cost = 1e30;
  const int cand[3] = {256, 128, 64};
  const double penalty[3] = {1.0, 1.12, 1.5};
  for (int i = 0; i < 3; ++i) {
    const long long nt = (n + cand[i] - 1) / cand[i];
    const long long waves = (m_tiles * nt + sms - 1) / sms;
    const double cost = double(waves) * cand[i] * penalty[i];
    if (cost < best_cost) {
      best_cost = cost;
      best = cand[i];
    }
  }
  return best;
}

}  // namespace mb200

using namespace mb200;

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
  MB_REQUIRE(a->k_splits >= 1, "mb_gemm: k_splits must be >= 1");
  MB_REQUIRE(a->k_splits == 1 || (a->epilogue & MB_EPI_ATOMIC),
             "mb_gemm: k_splits > 1 requires MB_EPI_ATOMIC");
  MB_REQUIRE(!(a->epilogue & MB_EPI_ATOMIC) || a->out_dtype == MB_F32,
             "mb_gemm: MB_EPI_ATOMIC requires an f32 output");
  MB_REQUIRE(!(a->epilogue & MB_EPI_DGELU) || a->aux_in, "mb_gemm: MB_EPI_DGELU needs aux_in");
  if (a->residual) MB_REQUIRE(a->ld_res % 4 == 0, "mb_gemm: ld_res must be a multiple of 4");
  if (a->aux_in || a->aux_out) MB_REQUIRE(a->ld_aux % 8 == 0, "mb_gemm: ld_aux must be a multiple of 8");
  const bool tf32 = (a->in_dtype == MB_F32);
  const int esize = tf32 ? 4 : 2;
  const int bk = 128 / esize;

  int bn = a->block_n ? a->block_n : pick_block_n(a->m, a->n);
  MB_REQUIRE(bn == 64 || bn == 128 || bn == 256, "mb_gemm: block_n %d unsupported", bn);

  GemmDev p;
  p.out = a->out;
  p.bias = a->bias;
  p.residual = a->residual;
  p.aux_in = reinterpret_cast<const __nv_bfloat16*>(a->aux_in);
  p.aux_out = reinterpret_cast<__nv_bfloat16*>(a->aux_out);
  p.M = (int)a->m;
  p.N = (int)a->n;
  p.K = (int)a->k;
  p.ldc = a->ldc;
  p.ld_res = a->ld_res;
  p.ld_aux = a->ld_aux;
  p.res_period = (int)a->res_period;
  p.out_f32 = (a->out_dtype == MB_F32);
  p.epilogue = a->epilogue;
  p.m_tiles = (p.M + kBM - 1) / kBM;
  p.n_tiles = (p.N + bn - 1) / bn;
  const int k_per_block = (a->a_layout == MB_MAJOR_MN || a->b_layout == MB_MAJOR_MN) ? 64 : bk;
  p.k_blocks = (p.K + k_per_block - 1) / k_per_block;
  p.k_splits = a->k_splits > p.k_blocks ? p.k_blocks : a->k_splits;
  p.kb_per_split = (p.k_blocks + p.k_splits - 1) / p.k_splits;
  p.k_splits = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.total_tiles = p.m_tiles * p.n_tiles * p.k_splits;
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

  CUtensorMap ta, tb;
  const TmaDtype tdt = tf32 ? kTmaF32 : kTmaBF16;
  // ---- A
  if (a->a_layout == MB_MAJOR_K) {
    uint64_t dims[2] = {(uint64_t)a->k, (uint64_t)a->m};
    uint64_t str[1] = {(uint64_t)a->lda * esize};
    uint32_t box[2] = {(uint32_t)bk, (uint32_t)kBM};
    if (make_tensor_map(&ta, a->a, tdt, 2, dims, str, box)) return -1;
  } else if (a->a_layout == MB_MAJOR_MN) {
    MB_REQUIRE(!tf32, "mb_gemm: MN-major A requires bf16");
    uint64_t dims[2] = {(uint64_t)a->m, (uint64_t)a->k};
    uint64_t str[1] = {(uint64_t)a->lda * esize};
    uint32_t box[2] = {64u, 64u};
    if (make_tensor_map(&ta, a->a, tdt, 2, dims, str, box)) return -1;
  } else if (a->a_layout == MB_A_PATCH32) {
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
  } else {
    MB_REQUIRE(false, "mb_gemm: bad a_layout %d", a->a_layout);
  }
  // ---- B
  if (a->b_layout == MB_MAJOR_K) {
    uint64_t dims[2] = {(uint64_t)a->k, (uint64_t)a->n};
    uint64_t str[1] = {(uint64_t)a->ldb * esize};
    uint32_t box[2] = {(uint32_t)bk, (uint32_t)bn};
    if (make_tensor_map(&tb, a->b, tdt, 2, dims, str, box)) return -1;
  } else if (a->b_layout == MB_MAJOR_MN) {
    MB_REQUIRE(!tf32, "mb_gemm: MN-major B requires bf16");
    uint64_t dims[2] = {(uint64_t)a->n, (uint64_t)a->k};
    uint64_t str[1] = {(uint64_t)a->ldb * esize};
    uint32_t box[2] = {64u, 64u};
    if (make_tensor_map(&tb, a->b, tdt, 2, dims, str, box)) return -1;
  } else {
    MB_REQUIRE(false, "mb_gemm: bad b_layout %d", a->b_layout);
  }

#define MB_LAUNCH(BN_, AL_, BMN_, ES_)                                   \
  if (bn == BN_) return launch_gemm<BN_, AL_, BMN_, ES_>(ta, tb, p, stream)

  if (a->a_layout == MB_MAJOR_K && a->b_layout == MB_MAJOR_K && !tf32) {
    MB_LAUNCH(256, MB_MAJOR_K, 0, 2);
    MB_LAUNCH(128, MB_MAJOR_K, 0, 2);
    MB_LAUNCH(64, MB_MAJOR_K, 0, 2);
  } else if (a->a_layout == MB_MAJOR_K && a->b_layout == MB_MAJOR_K && tf32) {
    MB_LAUNCH(256, MB_MAJOR_K, 0, 4);
    MB_LAUNCH(128, MB_MAJOR_K, 0, 4);
    MB_LAUNCH(64, MB_MAJOR_K, 0, 4);
  } else if (a->a_layout == MB_MAJOR_K && a->b_layout == MB_MAJOR_MN) {
    MB_LAUNCH(256, MB_MAJOR_K, 1, 2);
    MB_LAUNCH(128, MB_MAJOR_K, 1, 2);
    MB_LAUNCH(64, MB_MAJOR_K, 1, 2);
  } else if (a->a_layout == MB_MAJOR_MN && a->b_layout == MB_MAJOR_MN) {
    MB_LAUNCH(256, MB_MAJOR_MN, 1, 2);
    MB_LAUNCH(128, MB_MAJOR_MN, 1, 2);
    MB_LAUNCH(64, MB_MAJOR_MN, 1, 2);
  } else if (a->a_layout == MB_A_PATCH32 && a->b_layout == MB_MAJOR_K) {
    MB_LAUNCH(256, MB_A_PATCH32, 0, 4);
    MB_LAUNCH(128, MB_A_PATCH32, 0, 4);
    MB_LAUNCH(64, MB_A_PATCH32, 0, 4);
  }
#undef MB_LAUNCH
  MB_REQUIRE(false, "mb_gemm: unsupported layout combination a=%d b=%d dtype=%d", a->a_layout,
             a->b_layout, a->in_dtype);
}
