// mirage_b200/csrc/attention_small.cu
//
// Attention forward for SHORT sequences, head_dim 64, Nq <= 128 and Nk <= 128: the 98 visible tokens + 1 global
// token of a MultiMAE pretraining step (mirage/model.py:388-391 feeding Attention.forward, mirage/utils.py:181-185;
// BASELINE cfg 3 / 4: 256 x 16 (batch, head) problems of 99 x 99 per layer).
//
// Why a third kernel.  One (batch, head) problem is a single 128 x 128 score tile -- 3.3 MFLOP against 50 KB of
// operands: the layer is bound by HBM (Q, K, V read once, O written once: 207 MB = 32 us at cfg 4) and by the MUFU
// (128 x 128 exponentials = 1024 clk per problem and SM), not by the tensor core.  attention.cu walks such a
// problem through its two-tile flash pipeline with one of its two softmax groups idle and the online-softmax
// machinery (running max, rescale, statistics hand-off to a separate epilogue warpgroup) in the way: 96 us.
// Here a problem is ONE pass -- S = Q K^T, row max, exponentials, P V, normalise -- and an SM keeps FOUR problems
// in flight, each in its own slot:
//
//   slot s         smem: Q, K, V tiles (3 x 16 KB, TMA, 128-byte swizzle; rows >= N are zero-filled by TMA)
//                  TMEM: columns [128 s, 128 s + 128): S (fp32) -> P (bf16, written in place over columns [0, 64))
//                        -> O (fp32, columns [64, 128): free once every thread has read its S row)
//   warp 0         TMA producer: problem t goes to slot t % 4 as soon as S(t-4) has consumed Q/K and P V(t-4) V
//   warp 1         MMA issuer (whole warp, one elected lane): S(t), then P V(t-3) -- three problems are being
//                  exponentiated while the fourth drains
//   warps 4-19     warpgroup s owns slot s, thread r owns row r: two passes over its S row in 32-column chunks
//                  (max, then exp / sum / pack / store P), waits for P V, reads O, RELEASES the slot, then scales
//                  by 1 / l and stores bf16 O and the log-sum-exp.  No running max, no rescale, no shared memory.
//
// Columns >= Nk (zero-filled keys and the part of the tile the N = ceil16(Nk) score MMA does not write) are masked
// to -inf in both passes; P V runs over ceil16(Nk) keys only.
#include "attention_common.cuh"

namespace mb200 {

constexpr int kASlots = 4;
constexpr int kAThreads = 128 + kASlots * 128;   // control warpgroup (TMA, MMA, 2 idle) + 4 softmax warpgroups
constexpr int kALag = 3;                         // P V(t - kALag) is issued behind S(t)

struct ASCfg {
  static constexpr int kTile = 128 * 128;        // 128 rows x 64 bf16
  static constexpr int kSlot = 3 * kTile;        // Q, K, V
  static constexpr int kOffBar = kASlots * kSlot;
  static constexpr int kSmemBytes = kOffBar + 512;
};
static_assert(ASCfg::kSmemBytes <= 227 * 1024, "attention_small shared memory exceeds the 227 KB per-CTA limit");

__global__ void __launch_bounds__(kAThreads, 1)
attn_fwd_small_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const AttnDev p) {
  using Cfg = ASCfg;
  constexpr int HD = 64;
  constexpr uint64_t kSw = kDescSwizzle128B;
  constexpr uint32_t kSbo = 8 * 128;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* qk_full = bars;                      // TMA tx: Q and K of the slot's problem
  uint64_t* qk_free = bars + kASlots;            // MMA commit: S read them
  uint64_t* v_full = bars + 2 * kASlots;         // TMA tx
  uint64_t* v_free = bars + 3 * kASlots;         // MMA commit: P V read it
  uint64_t* s_full = bars + 4 * kASlots;         // MMA commit: S complete
  uint64_t* p_full = bars + 5 * kASlots;         // 4 softmax warps: P in TMEM
  uint64_t* o_full = bars + 6 * kASlots;         // MMA commit: P V complete
  uint64_t* o_free = bars + 7 * kASlots;         // 4 softmax warps: O in registers, slot reusable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 * kASlots);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("mirage_b200: attention_small smem base not 1024-byte aligned\n");
    __trap();
  }

  const int n_items = p.B * p.H;
  const int n_local = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                      static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    for (int s = 0; s < kASlots; ++s) {
      mbar_init(&qk_full[s], 1);
      mbar_init(&qk_free[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_free[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);
      mbar_init(&o_full[s], 1);
      mbar_init(&o_free[s], 4);
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
  pdl_wait();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // -------------------------------------------------------------- TMA producer (whole warp, uniform values)
      const bool leader = elect_one();
      for (int t = 0; t < n_local; ++t) {
        const int seq = static_cast<int>(blockIdx.x) + t * static_cast<int>(gridDim.x);
        const int item = p.reverse ? n_items - 1 - seq : seq;
        const int h = item % p.H, b = item / p.H;
        const int s = t & (kASlots - 1);
        const uint32_t ph = (t / kASlots) & 1;
        uint8_t* slot = smem + s * Cfg::kSlot;
        mbar_wait(&qk_free[s], ph ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&qk_full[s], 2 * Cfg::kTile);
          tma_load_3d(slot, &tm_q, &qk_full[s], h * HD, 0, b);
          tma_load_3d(slot + Cfg::kTile, &tm_k, &qk_full[s], h * HD, 0, b);
        }
        mbar_wait(&v_free[s], ph ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&v_full[s], Cfg::kTile);
          tma_load_3d(slot + 2 * Cfg::kTile, &tm_v, &v_full[s], h * HD, 0, b);
        }
        __syncwarp();
      }
      pdl_trigger();
    } else if (warp == 1) {
      // -------------------------------------------------------------- MMA issuer (whole warp, one elected lane)
      const bool leader = elect_one();
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t base = smem_u32(smem);
      const uint32_t nk16 = static_cast<uint32_t>((p.Nk + 15) & ~15);
      const uint32_t idesc_s = make_idesc(128, nk16, kFmtBF16, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(128, HD, kFmtBF16, 0, 1);
      const int ksteps = static_cast<int>(nk16 >> 4);
      for (int t = 0; t < n_local + kALag; ++t) {
        if (t < n_local) {
          const int s = t & (kASlots - 1);
          const uint32_t ph = (t / kASlots) & 1;
          mbar_wait(&qk_full[s], ph);
          mbar_wait(&o_free[s], ph ^ 1);          // the slot's previous O has been read out
          tc_fence_after();
          const uint64_t qd = make_smem_desc(base + s * Cfg::kSlot, 0, kSbo, kSw);
          const uint64_t kd = make_smem_desc(base + s * Cfg::kSlot + Cfg::kTile, 0, kSbo, kSw);
          if (leader) {
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)     // +32 bytes per K step = +2 in the descriptor's address field
              umma_f16_ss(tb + s * 128, qd + 2 * k, kd + 2 * k, idesc_s, k > 0 ? 1u : 0u);
            umma_commit(&s_full[s]);
            umma_commit(&qk_free[s]);
          }
          __syncwarp();
        }
        const int u = t - kALag;
        if (u >= 0) {
          const int s = u & (kASlots - 1);
          const uint32_t ph = (u / kASlots) & 1;
          mbar_wait(&v_full[s], ph);
          mbar_wait(&p_full[s], ph);
          tc_fence_after();
          const uint64_t vd = make_smem_desc(base + s * Cfg::kSlot + 2 * Cfg::kTile, 0, kSbo, kSw);
          if (leader) {
            for (int kk = 0; kk < ksteps; ++kk)   // 16 keys = 2048 bytes = +128 in the address field
              umma_f16_ts(tb + s * 128 + 64, tb + s * 128 + kk * 8, vd + 128 * kk, idesc_pv, kk > 0 ? 1u : 0u);
            umma_commit(&o_full[s]);
            umma_commit(&v_free[s]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue: warpgroup = slot, thread = row
    // register pool of the CTA = 640 x 96 at launch = 61440 >= 128 x (40 + 4 x 104)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int s = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                     // row of the tile == TMEM lane
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + s * 128;
    const uint32_t t_o = t_s + 64;
    const int nk = p.Nk;
    const int nchunks = (nk + 31) >> 5;

    for (int t = s; t < n_local; t += kASlots) {
      const int seq = static_cast<int>(blockIdx.x) + t * static_cast<int>(gridDim.x);
      const int item = p.reverse ? n_items - 1 - seq : seq;
      const int h = item % p.H, b = item / p.H;
      const uint32_t ph = (t / kASlots) & 1;
      mbar_wait(&s_full[s], ph);
      tc_fence_after();

      // pass 1: row maximum
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nchunks) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_s + c * 32, v);
          tmem_ld_wait();
          if (c * 32 + 32 <= nk) {
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[i]));
            mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < nk) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        }
      }
      const float neg_m = -mx * p.scale_log2;              // scale > 0

      // pass 2: e = 2^(s * scale_log2 - m), row sum, bf16 P over the first half of S
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nchunks) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_s + c * 32, v);
          tmem_ld_wait();
          const bool full = (c * 32 + 32 <= nk);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float e = fast_exp2(fmaf(__uint_as_float(v[i]), p.scale_log2, neg_m));
            if (!full && c * 32 + i >= nk) e = 0.f;
            l4[i & 3] += e;
            v[i] = __float_as_uint(e);
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
          tmem_st_32x32b_x16(t_s + c * 16, pk);            // columns < 16 (c + 1) <= 32 c: S chunks already consumed
        }
      }
      const float l_row = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[s]);

      // epilogue: O to registers, release the slot, then normalise and store
      mbar_wait(&o_full[s], ph);
      tc_fence_after();
      uint32_t o[64];
      tmem_ld_32x32b_x32_p(t_o, o);
      tmem_ld_32x32b_x32_p(t_o + 32, o + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[s]);
      if (r < p.Nq) {
        const float inv_l = 1.f / l_row;
        __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.Nq + r) * p.ldo + h * HD;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(o[c * 8 + 0]) * inv_l, __uint_as_float(o[c * 8 + 1]) * inv_l);
          pk.y = pack_bf16x2(__uint_as_float(o[c * 8 + 2]) * inv_l, __uint_as_float(o[c * 8 + 3]) * inv_l);
          pk.z = pack_bf16x2(__uint_as_float(o[c * 8 + 4]) * inv_l, __uint_as_float(o[c * 8 + 5]) * inv_l);
          pk.w = pack_bf16x2(__uint_as_float(o[c * 8 + 6]) * inv_l, __uint_as_float(o[c * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 8) = pk;
        }
        if (p.lse != nullptr)
          p.lse[(static_cast<long long>(b) * p.H + h) * p.Nq + r] =
              (-neg_m + log2f(l_row)) * 0.6931471805599453f;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// Fits: head_dim 64, one query tile, one key tile.  Returns 1 when the problem does not fit, < 0 on error.
int launch_attn_fwd_small(const mb_attn_args* a, const AttnDev& p_in, cudaStream_t stream) {
  if (a->head_dim != 64 || a->nq > 128 || a->nk > 128) return 1;
  AttnDev p = p_in;
  p.reverse = take_direction() < 0 ? 1 : 0;
  CUtensorMap tq, tk, tv;
  {
    uint64_t dims[3] = {(uint64_t)a->heads * 64, (uint64_t)a->nq, (uint64_t)a->batch};
    uint64_t str[2] = {(uint64_t)a->ldq * 2, (uint64_t)a->nq * a->ldq * 2};
    uint32_t box[3] = {64u, 128u, 1u};
    if (make_tensor_map(&tq, a->q, kTmaBF16, 3, dims, str, box, 128)) return -1;
  }
  {
    uint64_t dims[3] = {(uint64_t)a->heads * 64, (uint64_t)a->nk, (uint64_t)a->batch};
    uint64_t str[2] = {(uint64_t)a->ldk * 2, (uint64_t)a->nk * a->ldk * 2};
    uint32_t box[3] = {64u, 128u, 1u};
    if (make_tensor_map(&tk, a->k, kTmaBF16, 3, dims, str, box, 128)) return -1;
    uint64_t strv[2] = {(uint64_t)a->ldv * 2, (uint64_t)a->nk * a->ldv * 2};
    if (make_tensor_map(&tv, a->v, kTmaBF16, 3, dims, strv, box, 128)) return -1;
  }
  auto kern = attn_fwd_small_kernel;
  static PerDeviceOnce configured;
  if (configured.first())
    MB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ASCfg::kSmemBytes));
  const long long items = static_cast<long long>(p.B) * p.H;
  MB_REQUIRE(items > 0 && items < (1ll << 31), "mb_attn_fwd: %lld work items out of range", items);
  const long long grid = items < sm_count() ? items : sm_count();
  MB_CHECK_CUDA(launch_k(kern, dim3(static_cast<unsigned>(grid)), dim3(kAThreads), ASCfg::kSmemBytes, stream, tq, tk, tv,
                         p));
  return 0;
}

}  // namespace mb200
