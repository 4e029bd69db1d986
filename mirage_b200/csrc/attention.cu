// mirage_b200/csrc/attention.cu
//
// Fused flash-style attention forward on tcgen05/TMEM for sm_100a (no mask, no dropout):
//     O = softmax(Q K^T * scale) V        per (batch, head)
// Replaces F.scaled_dot_product_attention + the transpose/reshape copy at
//   mirage/utils.py:181-185 (Attention)  and  mirage/utils.py:216-220 (CrossAttention).
//
// One CTA owns up to TWO 128-row query tiles of one (batch, head) and walks the keys in blocks of
// 128 (the last block is shortened to a multiple of 16 keys).
//
//   warp 0      TMA producer: Q tiles once, then K_j / V_j through a 3-stage ring
//   warp 1      TMEM allocator + MMA issuer (one lane):
//                   S_g = Q_g K_j^T      (SS, both K-major,      accumulator S_g in TMEM)
//                   O_g += P_g V_j       (SS, P K-major from smem, V MN-major, accumulator O_g)
//   warps 2-5   softmax group 0 (one thread per query row, no shuffles)
//   warps 6-9   softmax group 1
//
// The two groups ping-pong: while group 0 exponentiates S_0(j+1) the tensor core runs P_1 V_j and
// S_1(j+1).  Softmax runs in the log2 domain with a running max m and sum l per row; O is rescaled
// in TMEM (tcgen05.ld / st) only when some row max in the warp actually moved.
//
// TMEM columns: S_0 [0,128)  S_1 [128,256)  O_0 [256,256+HD)  O_1 [320,320+HD).
#include "../../include/mirage_b200.h"
#include "common.cuh"

namespace mb200 {

struct AttnDev {
  __nv_bfloat16* out;
  float* lse;
  long long ldo;
  int B, H, Nq, Nk;
  int q_tiles, q_pairs, kv_blocks;
  float scale_log2;
};

constexpr int kAttnThreads = 320;
constexpr int kKvStages = 3;

template <int HD>
struct AttnCfg {
  static constexpr int kRowBytes = HD * 2;          // 128 (SW128) or 64 (SW64)
  static constexpr int kTileBytes = 128 * kRowBytes;  // Q / K / V tile of 128 rows
  static constexpr int kPBytes = 128 * 128 * 2;     // P tile: 128 rows x 128 keys, bf16
  static constexpr int kOffQ = 0;
  static constexpr int kOffK = kOffQ + 2 * kTileBytes;
  static constexpr int kOffV = kOffK + kKvStages * kTileBytes;
  static constexpr int kOffP = kOffV + kKvStages * kTileBytes;
  static constexpr int kOffBar = kOffP + 2 * kPBytes;
  static constexpr int kSmemBytes = kOffBar + 256;
  static constexpr int kSwizzle = (HD == 64) ? 128 : 64;
};

template <int HD>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const AttnDev p) {
  using Cfg = AttnCfg<HD>;
  constexpr uint64_t kSw = (HD == 64) ? kDescSwizzle128B : kDescSwizzle64B;
  constexpr uint32_t kSbo = 8 * Cfg::kRowBytes;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* q_full = bars;                    // 1
  uint64_t* k_full = bars + 1;                // 3
  uint64_t* k_empty = bars + 4;               // 3
  uint64_t* v_full = bars + 7;                // 3
  uint64_t* v_empty = bars + 10;              // 3
  uint64_t* s_full = bars + 13;               // 2
  uint64_t* p_full = bars + 15;               // 2
  uint64_t* o_full = bars + 17;               // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("mirage_b200: attention smem base not 1024-byte aligned\n");
    __trap();
  }

  // work decode
  const int pair = blockIdx.x % p.q_pairs;
  const int bh = blockIdx.x / p.q_pairs;
  const int h = bh % p.H;
  const int b = bh / p.H;
  const int n_groups = (pair * 2 + 1 < p.q_tiles) ? 2 : 1;
  const int kvb = p.kv_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 128);
      mbar_init(&o_full[g], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ---------------------------------------------------------------- TMA producer
    mbar_arrive_expect_tx(q_full, n_groups * Cfg::kTileBytes);
    for (int g = 0; g < n_groups; ++g)
      tma_load_3d(smem + Cfg::kOffQ + g * Cfg::kTileBytes, &tm_q, q_full, h * HD,
                  (pair * 2 + g) * 128, b);
    for (int j = 0; j < kvb; ++j) {
      const int s = j % kKvStages;
      const uint32_t ph = (j / kKvStages) & 1;
      mbar_wait(&k_empty[s], ph ^ 1);
      mbar_arrive_expect_tx(&k_full[s], Cfg::kTileBytes);
      tma_load_3d(smem + Cfg::kOffK + s * Cfg::kTileBytes, &tm_k, &k_full[s], h * HD, j * 128, b);
      mbar_wait(&v_empty[s], ph ^ 1);
      mbar_arrive_expect_tx(&v_full[s], Cfg::kTileBytes);
      tma_load_3d(smem + Cfg::kOffV + s * Cfg::kTileBytes, &tm_v, &v_full[s], h * HD, j * 128, b);
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t q_addr = smem_u32(smem + Cfg::kOffQ);
    const uint32_t k_addr = smem_u32(smem + Cfg::kOffK);
    const uint32_t v_addr = smem_u32(smem + Cfg::kOffV);
    const uint32_t p_addr = smem_u32(smem + Cfg::kOffP);
    constexpr uint32_t idesc_pv = make_idesc(128, HD, kFmtBF16, 0, 1);

    auto issue_s = [&](int g, int j) {
      const int s = j % kKvStages;
      const int valid = min(128, p.Nk - j * 128);
      const uint32_t ncols = static_cast<uint32_t>((valid + 15) & ~15);
      const uint32_t idesc_s = make_idesc(128, ncols, kFmtBF16, 0, 0);
      const uint32_t qa = q_addr + g * Cfg::kTileBytes;
      const uint32_t ka = k_addr + s * Cfg::kTileBytes;
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        const uint64_t da = make_smem_desc(qa + k * 32, 0, kSbo, kSw);
        const uint64_t db = make_smem_desc(ka + k * 32, 0, kSbo, kSw);
        umma_f16_ss(tmem_base + g * 128, da, db, idesc_s, k > 0 ? 1u : 0u);
      }
      umma_commit(&s_full[g]);
      if (g == n_groups - 1) umma_commit(&k_empty[s]);
    };

    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    for (int g = 0; g < n_groups; ++g) issue_s(g, 0);

    for (int j = 0; j < kvb; ++j) {
      const int s = j % kKvStages;
      const int valid = min(128, p.Nk - j * 128);
      const int ksteps = (valid + 15) >> 4;
      for (int g = 0; g < n_groups; ++g) {
        mbar_wait(&p_full[g], j & 1);
        if (g == 0) mbar_wait(&v_full[s], (j / kKvStages) & 1);
        tc_fence_after();
        const uint32_t pa = p_addr + g * Cfg::kPBytes;
        const uint32_t va = v_addr + s * Cfg::kTileBytes;
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint64_t da = make_smem_desc(pa + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024);
          const uint64_t db = make_smem_desc(va + kk * 16 * Cfg::kRowBytes, 0, kSbo, kSw);
          umma_f16_ss(tmem_base + 256 + g * 64, da, db, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&o_full[g]);
        if (g == n_groups - 1) umma_commit(&v_empty[s]);
        if (j + 1 < kvb) {
          if (g == 0) {
            mbar_wait(&k_full[(j + 1) % kKvStages], ((j + 1) / kKvStages) & 1);
            tc_fence_after();
          }
          issue_s(g, j + 1);
        }
      }
    }
  } else if (warp >= 2) {
    // ---------------------------------------------------------------- softmax / epilogue
    const int g = (warp - 2) >> 2;
    if (g < n_groups) {
      const int quarter = warp & 3;
      const int r = quarter * 32 + lane;                    // row inside the tile == TMEM lane
      const int qrow = (pair * 2 + g) * 128 + r;            // query index inside this (b, h)
      const bool warp_live = (pair * 2 + g) * 128 + quarter * 32 < p.Nq;
      const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + g * 128;
      const uint32_t t_o = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + 256 + g * 64;
      uint8_t* p_row = smem + Cfg::kOffP + g * Cfg::kPBytes + r * 128;
      const uint32_t sw = static_cast<uint32_t>(r & 7);
      float m_run = -INFINITY;
      float l_run = 0.f;

      for (int j = 0; j < kvb; ++j) {
        const int valid = min(128, p.Nk - j * 128);
        const int nchunks = (valid + 31) >> 5;
        mbar_wait(&s_full[g], j & 1);
        tc_fence_after();
        float alpha = 1.f;
        float m_new = m_run;
        if (warp_live) {
          // pass 1: row max
          float mx = -INFINITY;
          for (int c = 0; c < nchunks; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_s + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
          m_new = fmaxf(m_run, mx * p.scale_log2);
          alpha = fast_exp2(m_run - m_new);
        }
        if (j > 0) {
          mbar_wait(&o_full[g], (j - 1) & 1);  // P_g and O_g are free again
          tc_fence_after();
          if (warp_live && __any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
            for (int c = 0; c < HD / 32; ++c) {
              uint32_t v[32];
              tmem_ld_32x32b_x32(t_o + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st_32x32b_x32(t_o + c * 32, v);
            }
            tmem_st_wait();
          }
        }
        if (warp_live) {
          l_run *= alpha;
          float sum = 0.f;
          for (int c = 0; c < nchunks; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_s + c * 32, v);
            tmem_ld_wait();
            float e[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float x = fast_exp2(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new));
              e[i] = (c * 32 + i < valid) ? x : 0.f;
              sum += e[i];
            }
            uint8_t* dst = p_row + (c >> 1) * 16384;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              uint4 pk;
              pk.x = pack_bf16x2(e[t * 8 + 0], e[t * 8 + 1]);
              pk.y = pack_bf16x2(e[t * 8 + 2], e[t * 8 + 3]);
              pk.z = pack_bf16x2(e[t * 8 + 4], e[t * 8 + 5]);
              pk.w = pack_bf16x2(e[t * 8 + 6], e[t * 8 + 7]);
              const uint32_t c8 = static_cast<uint32_t>((c & 1) * 4 + t);
              *reinterpret_cast<uint4*>(dst + ((c8 ^ sw) << 4)) = pk;
            }
          }
          l_run += sum;
          m_run = m_new;
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(&p_full[g]);
      }

      // epilogue: O / l -> bf16 -> global, one 2*HD-byte row per thread
      mbar_wait(&o_full[g], (kvb - 1) & 1);
      tc_fence_after();
      if (warp_live) {
        const float inv_l = 1.f / l_run;
        __nv_bfloat16* orow =
            p.out + (static_cast<long long>(b) * p.Nq + qrow) * p.ldo + h * HD;
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_o + c * 32, v);
          tmem_ld_wait();
          if (qrow < p.Nq) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              uint4 pk;
              pk.x = pack_bf16x2(__uint_as_float(v[t * 8 + 0]) * inv_l,
                                 __uint_as_float(v[t * 8 + 1]) * inv_l);
              pk.y = pack_bf16x2(__uint_as_float(v[t * 8 + 2]) * inv_l,
                                 __uint_as_float(v[t * 8 + 3]) * inv_l);
              pk.z = pack_bf16x2(__uint_as_float(v[t * 8 + 4]) * inv_l,
                                 __uint_as_float(v[t * 8 + 5]) * inv_l);
              pk.w = pack_bf16x2(__uint_as_float(v[t * 8 + 6]) * inv_l,
                                 __uint_as_float(v[t * 8 + 7]) * inv_l);
              *reinterpret_cast<uint4*>(orow + c * 32 + t * 8) = pk;
            }
          }
        }
        if (p.lse != nullptr && qrow < p.Nq)
          p.lse[(static_cast<long long>(b) * p.H + h) * p.Nq + qrow] =
              (m_run + log2f(l_run)) * 0.6931471805599453f;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int HD>
static int launch_attn_fwd(const mb_attn_args* a, cudaStream_t stream) {
  using Cfg = AttnCfg<HD>;
  CUtensorMap tq, tk, tv;
  {
    uint64_t dims[3] = {(uint64_t)a->heads * HD, (uint64_t)a->nq, (uint64_t)a->batch};
    uint64_t str[2] = {(uint64_t)a->ldq * 2, (uint64_t)a->nq * a->ldq * 2};
    uint32_t box[3] = {(uint32_t)HD, 128u, 1u};
    if (make_tensor_map(&tq, a->q, kTmaBF16, 3, dims, str, box, Cfg::kSwizzle)) return -1;
  }
  {
    uint64_t dims[3] = {(uint64_t)a->heads * HD, (uint64_t)a->nk, (uint64_t)a->batch};
    uint64_t str[2] = {(uint64_t)a->ldk * 2, (uint64_t)a->nk * a->ldk * 2};
    uint32_t box[3] = {(uint32_t)HD, 128u, 1u};
    if (make_tensor_map(&tk, a->k, kTmaBF16, 3, dims, str, box, Cfg::kSwizzle)) return -1;
    uint64_t strv[2] = {(uint64_t)a->ldv * 2, (uint64_t)a->nk * a->ldv * 2};
    if (make_tensor_map(&tv, a->v, kTmaBF16, 3, dims, strv, box, Cfg::kSwizzle)) return -1;
  }
  AttnDev p;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.lse = a->lse;
  p.ldo = a->ldo;
  p.B = (int)a->batch;
  p.H = (int)a->heads;
  p.Nq = (int)a->nq;
  p.Nk = (int)a->nk;
  p.q_tiles = (p.Nq + 127) / 128;
  p.q_pairs = (p.q_tiles + 1) / 2;
  p.kv_blocks = (p.Nk + 127) / 128;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  auto kern = attn_fwd_kernel<HD>;
  static bool configured = false;
  if (!configured) {
    MB_CHECK_CUDA(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const long long grid = (long long)p.B * p.H * p.q_pairs;
  MB_REQUIRE(grid > 0 && grid < (1ll << 31), "mb_attn_fwd: grid %lld out of range", grid);
  kern<<<(unsigned)grid, kAttnThreads, Cfg::kSmemBytes, stream>>>(tq, tk, tv, p);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mb200

using namespace mb200;

extern "C" int mb_attn_fwd(const mb_attn_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(a != nullptr, "mb_attn_fwd: null args");
  MB_REQUIRE(a->q && a->k && a->v && a->out, "mb_attn_fwd: null tensor pointer");
  MB_REQUIRE(a->batch > 0 && a->heads > 0 && a->nq > 0 && a->nk > 0, "mb_attn_fwd: empty problem");
  MB_REQUIRE(a->head_dim == 64 || a->head_dim == 32, "mb_attn_fwd: head_dim %d unsupported (32|64)",
             a->head_dim);
  MB_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0,
             "mb_attn_fwd: leading dimensions must be multiples of 8 elements");
  if (a->head_dim == 64) return launch_attn_fwd<64>(a, stream);
  return launch_attn_fwd<32>(a, stream);
}
