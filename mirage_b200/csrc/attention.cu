// mirage_b200/csrc/attention.cu
//
// Fused flash-style attention forward on tcgen05/TMEM for sm_100a (no mask, no dropout):
//     O = softmax(Q K^T * scale) V        per (batch, head)
// Replaces F.scaled_dot_product_attention + the transpose/reshape copy at
//   mirage/utils.py:181-185 (Attention)  and  mirage/utils.py:216-220 (CrossAttention).
//
// Persistent CTAs (one per SM) walk work items = (batch, head, pair of 128-row query tiles); within an
// item the keys are walked in blocks of 128 (a ragged last block is shortened to a multiple of 16).
// Every role runs ahead across item boundaries (Q is double-buffered), so there is no per-item
// pipeline fill or drain on the softmax warps:
//
//   warp 0       TMA producer: Q tiles (+ the peeled key/value tail rows) per item, K_j / V_j through
//                3-stage rings
//   warp 1       TMEM allocator + MMA issuer (one lane), S always one block ahead of P V:
//                    S_g = Q_g K_j^T      (SS: both operands K-major in smem, accumulator S_g in TMEM)
//                    O_g += P_g V_j       (TS: P_g read from TMEM as the A operand, V MN-major in smem)
//   warps 4-11   softmax group 0: TWO threads per query row (warps 4-7 own key columns [0,64) of the
//   warps 12-19  softmax group 1   block, warps 8-11 columns [64,128)); the two halves exchange their
//                row maxima through smem + a 64-thread named barrier (per 32 rows), everything else is thread-private.  Four
//                softmax warps per SM sub-partition instead of two: the first layout (one thread per
//                row, 128 values in registers) left the MUFU idle 2/3 of the time for lack of warps to
//                switch to (ncu r01 v5: issue 38 %, XU 36 %, top stall "wait")
//   warps 20-23  epilogue: O_g / l -> bf16 -> global (+ LSE) once the item's last P V retired, so the
//                softmax warps never wait for the drain of their own accumulator
//
// Each softmax thread pulls its half S row (64 fp32) into registers with one round of tcgen05.ld and
// releases the S buffer at once (s_free): S_g(j+1) runs while the row is being exponentiated.  The
// exponentials are packed to bf16 in registers and written back to TMEM (P_g) with tcgen05.st -- no
// shared-memory round trip, no proxy fence.  Softmax runs in the log2 domain with a running max m and
// sum l per row (packed FFMA2 / FADD2).  The max is LAZY: m only moves (and O is only rescaled in
// TMEM) when some row of the warp grew by more than 2^8 -- P stays <= 256, well inside bf16/fp32
// range, and the final O / l is unchanged -- so the TMEM read-modify-write of O is off the common path.
//
// TMEM columns: S_0 [0,128)  S_1 [128,256)  O_0 [256,256+HD)  O_1 [320,320+HD)  P_0 [384,448)  P_1 [448,512).
#include "../../include/mirage_b200.h"
#include <cstdlib>

#include "attention_common.cuh"

namespace mb200 {

template <int HD>
struct AttnCfg {
  static constexpr int kRowBytes = HD * 2;            // 128 (SW128) or 64 (SW64)
  static constexpr int kTileBytes = 128 * kRowBytes;  // Q / K / V tile of 128 rows
  static constexpr int kOffQ = 0;                     // 2 item slots x 2 tiles
  static constexpr int kOffK = kOffQ + 4 * kTileBytes;
  static constexpr int kOffV = kOffK + kKvStages * kTileBytes;
  static constexpr int kOffTail = kOffV + kKvStages * kTileBytes;  // 2 slots x {k rows, v rows}
  static constexpr int kTailSlotBytes = 2 * kMaxTail * 128;
  static constexpr int kOffStats = kOffTail + 2 * kTailSlotBytes;  // [parity][group][field][row] f32
  static constexpr int kStatFields = 3 + kMaxTail;                 // l (half 0), l (half 1), m, e_tail[]
  static constexpr int kStatsBytes = 2 * 2 * kStatFields * 128 * 4;
  static constexpr int kOffXchg = kOffStats + kStatsBytes;         // [parity][group][half][row] f32 row maxima
  static constexpr int kOffBar = kOffXchg + 2 * 2 * 2 * 128 * 4;
  static constexpr int kSmemBytes = kOffBar + 512;
  static constexpr int kSwizzle = (HD == 64) ? 128 : 64;
};

template <int HD, int POLY>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const AttnDev p) {
  using Cfg = AttnCfg<HD>;
  constexpr uint64_t kSw = (HD == 64) ? kDescSwizzle128B : kDescSwizzle64B;
  constexpr uint32_t kSbo = 8 * Cfg::kRowBytes;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* q_full = bars;                    // 2   TMA tx (Q tiles + tail rows of the item)
  uint64_t* q_empty = bars + 2;               // 2   MMA commit after the item's last S
  uint64_t* k_full = bars + 4;                // 3
  uint64_t* k_empty = bars + 7;               // 3
  uint64_t* v_full = bars + 10;               // 3
  uint64_t* v_empty = bars + 13;              // 3
  uint64_t* s_full = bars + 16;               // 2   MMA commit: S_g(j) complete
  uint64_t* s_free = bars + 18;               // 2   256 softmax threads: S_g copied to registers
  uint64_t* p_full = bars + 20;               // 2   256 softmax threads: P_g(j) in TMEM
  uint64_t* o_full = bars + 22;               // 2   MMA commit: P_g V(j) retired
  uint64_t* o_free = bars + 24;               // 2   4 epilogue warps: O_g of the item read out
  uint64_t* stats_full = bars + 26;           // 2   256 softmax threads: row statistics of the item
  uint64_t* tail_free = bars + 28;            // 2   4 epilogue warps: tail rows of the item slot consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);
  float* stats = reinterpret_cast<float*>(smem + Cfg::kOffStats);
  float* xchg = reinterpret_cast<float*>(smem + Cfg::kOffXchg);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("mirage_b200: attention smem base not 1024-byte aligned\n");
    __trap();
  }

  const int kvb = p.kv_blocks;
  const int n_items = p.B * p.H * p.q_pairs;  // work item = (batch, head, pair of query tiles)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 2);  // one commit per MMA issuer
      mbar_init(&tail_free[s], 4);
    }
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 2);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 2);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&s_free[g], 8);     // 8 softmax warps of the group, one elected arrival each
      mbar_init(&p_full[g], 8);
      mbar_init(&o_full[g], 1);
      mbar_init(&o_free[g], 4);
      mbar_init(&stats_full[g], 8);
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

  auto groups_of = [&](int item) { return ((item % p.q_pairs) * 2 + 1 < p.q_tiles) ? 2 : 1; };

  // All barrier phases are tracked with running counters; every role walks the same static list of
  // work items.
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    // (producer and issuers: whole converged warps, one elected lane around the TMA / MMA instructions -- see attention4.cu)
    if (warp == 0) {
      // -------------------------------------------------------------- TMA producer
      const bool leader = elect_one();
      int kv_count = 0, item_i = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_i) {
        const int pair = item % p.q_pairs;
        const int bh = item / p.q_pairs;
        const int h = bh % p.H, b = bh / p.H;
        const int n_groups = groups_of(item);
        const int slot = item_i & 1;
        const uint32_t ph = (item_i >> 1) & 1;
        mbar_wait(&q_empty[slot], ph ^ 1);
        uint32_t bytes = n_groups * Cfg::kTileBytes;
        if (HD == 64 && p.k_tail > 0) {
          mbar_wait(&tail_free[slot], ph ^ 1);
          bytes += 2 * p.k_tail * 128;
        }
        if (leader) {
        mbar_arrive_expect_tx(&q_full[slot], bytes);
        for (int g = 0; g < n_groups; ++g)
          tma_load_3d(smem + Cfg::kOffQ + (slot * 2 + g) * Cfg::kTileBytes, &tm_q, &q_full[slot], h * HD,
                      (pair * 2 + g) * 128, b);
        if (HD == 64) {
          uint8_t* tslot = smem + Cfg::kOffTail + slot * Cfg::kTailSlotBytes;
          for (int t = 0; t < p.k_tail; ++t) {
            const long long row = static_cast<long long>(b) * p.Nk + p.Nk_main + t;
            bulk_load(tslot + t * 128, p.k + row * p.ldk + h * HD, 128, &q_full[slot]);
            bulk_load(tslot + (kMaxTail + t) * 128, p.v + row * p.ldv + h * HD, 128, &q_full[slot]);
          }
        }
        }
        __syncwarp();
        for (int j = 0; j < kvb; ++j, ++kv_count) {
          const int s = kv_count % kKvStages;
          const uint32_t kph = (kv_count / kKvStages) & 1;
          mbar_wait(&k_empty[s], kph ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(&k_full[s], Cfg::kTileBytes);
            tma_load_3d(smem + Cfg::kOffK + s * Cfg::kTileBytes, &tm_k, &k_full[s], h * HD, j * 128, b);
          }
          mbar_wait(&v_empty[s], kph ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(&v_full[s], Cfg::kTileBytes);
            tma_load_3d(smem + Cfg::kOffV + s * Cfg::kTileBytes, &tm_v, &v_full[s], h * HD, j * 128, b);
          }
          __syncwarp();
        }
      }
      pdl_trigger();
    } else if (warp == 1 || warp == 2) {
      // -------------------------------------------------------------- MMA issuers
      // One issuing thread PER softmax group (warp 1 -> group 0, warp 2 -> group 1), each with blocking
      // waits in its own group's event order: S_g(j+1), then P_g V(j).  S runs one key block ahead of
      // P V -- also across item boundaries (the next item's Q sits in the other slot).  A single
      // in-order issuer coupled the groups: a wait for one group's P delayed the other group's S and
      // pulled both groups into lockstep, where they fight for the MUFU at the same time.
      // The shared K / V / Q slots are released by two commits (one per issuer; the issuer of group 0
      // commits twice for an item that has only one query tile).
      const int g = __shfl_sync(0xffffffffu, warp, 0) - 1;
      const bool leader = elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t q_addr = smem_u32(smem + Cfg::kOffQ);
      const uint32_t k_addr = smem_u32(smem + Cfg::kOffK);
      const uint32_t v_addr = smem_u32(smem + Cfg::kOffV);
      constexpr uint32_t idesc_pv = make_idesc(128, HD, kFmtBF16, 0, 1);
      int sc = 0;  // S MMAs issued        (phase of s_full / s_free)
      int pc = 0;  // P V MMAs issued      (phase of p_full / o_full)
      int oc = 0;  // items finished       (phase of o_free)

      auto issue_s = [&](int n_groups, int item_i, int j, int kc) {
        if (sc > 0) mbar_wait(&s_free[g], (sc - 1) & 1);
        mbar_wait(&k_full[kc % kKvStages], (kc / kKvStages) & 1);
        tc_fence_after();
        const int valid = min(128, p.Nk_main - j * 128);
        const uint32_t idesc_s = make_idesc(128, static_cast<uint32_t>((valid + 15) & ~15), kFmtBF16, 0, 0);
        const uint32_t qa = q_addr + ((item_i & 1) * 2 + g) * Cfg::kTileBytes;
        const uint32_t ka = k_addr + (kc % kKvStages) * Cfg::kTileBytes;
        const uint64_t qd = make_smem_desc(qa, 0, kSbo, kSw), kd = make_smem_desc(ka, 0, kSbo, kSw);
        if (leader) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)   // +32 bytes per K step = +2 in the 16-byte address field
            umma_f16_ss(tmem_u + g * 128, qd + 2 * k, kd + 2 * k, idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&s_full[g]);
          umma_commit(&k_empty[kc % kKvStages]);
          if (n_groups == 1) umma_commit(&k_empty[kc % kKvStages]);
          if (j == kvb - 1) {  // last S of the item: its Q slot may go
            umma_commit(&q_empty[item_i & 1]);
            if (n_groups == 1) umma_commit(&q_empty[item_i & 1]);
          }
        }
        __syncwarp();
        ++sc;
      };
      auto issue_pv = [&](int n_groups, int j, int kc) {
        mbar_wait(&p_full[g], pc & 1);
        if (j == 0 && oc > 0) mbar_wait(&o_free[g], (oc - 1) & 1);  // previous item's O read out
        mbar_wait(&v_full[kc % kKvStages], (kc / kKvStages) & 1);
        tc_fence_after();
        const int valid = min(128, p.Nk_main - j * 128);
        const int ksteps = (valid + 15) >> 4;
        const uint32_t va = v_addr + (kc % kKvStages) * Cfg::kTileBytes;
        const uint64_t vd = make_smem_desc(va, 0, kSbo, kSw);
        if (leader) {
          for (int kk = 0; kk < ksteps; ++kk)   // 16 keys = 16 rows of the V tile
            umma_f16_ts(tmem_u + 256 + g * 64, tmem_u + 384 + g * 64 + kk * 8, vd + kk * (Cfg::kRowBytes),
                        idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&o_full[g]);
          umma_commit(&v_empty[kc % kKvStages]);
          if (n_groups == 1) umma_commit(&v_empty[kc % kKvStages]);
        }
        __syncwarp();
        ++pc;
        if (j == kvb - 1) ++oc;
      };

      int kv_base = 0, item_i = 0;
      bool s0_issued = false;  // S(item, 0) already issued as the previous item's look-ahead
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_i, kv_base += kvb) {
        const int n_groups = groups_of(item);
        if (g >= n_groups) continue;  // the other issuer releases this item's slots for both
        if (!s0_issued) {
          mbar_wait(&q_full[item_i & 1], (item_i >> 1) & 1);
          issue_s(n_groups, item_i, 0, kv_base);
        }
        s0_issued = false;
        const int next = item + gridDim.x;
        for (int j = 0; j < kvb; ++j) {
          if (j + 1 < kvb) {
            issue_s(n_groups, item_i, j + 1, kv_base + j + 1);
          } else if (next < n_items && g < groups_of(next)) {
            mbar_wait(&q_full[(item_i + 1) & 1], ((item_i + 1) >> 1) & 1);
            issue_s(groups_of(next), item_i + 1, 0, kv_base + kvb);
            s0_issued = true;
          }
          issue_pv(n_groups, j, kv_base + j);
        }
      }
    }
  } else if (warp < 20) {
    // ---------------------------------------------------------------- softmax
    // register pool of the CTA = 768 threads x 80 at launch = 61440 = 128 x (48 + 4 x 96 + 48)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    const int sgi = (warp - 4) >> 2;
    const int g = sgi >> 1;        // query tile of the item
    const int hf = sgi & 1;        // key columns [hf*64, hf*64+64) of every block
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                    // row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + lane_off + g * 128 + hf * 64;
    const uint32_t t_o = tmem_base + lane_off + 256 + g * 64;
    const uint32_t t_p = tmem_base + lane_off + 384 + g * 64 + hf * 32;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    int cnt = 0;   // key blocks this group has processed (phases of s_full / o_full)
    int ic = 0;    // items this group has processed     (stats slot / phases of stats_full, o_free)
    int item_i = 0;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_i) {
      const int pair = item % p.q_pairs;
      if (pair * 2 + g >= p.q_tiles) continue;            // this group has no tile in the item
      const bool warp_live = (pair * 2 + g) * 128 + quarter * 32 < p.Nq_main;
      float m_run = -INFINITY;
      float l_run = 0.f;  // sum over THIS half's columns; the halves are added by the epilogue

      // Peeled key tail (HD == 64, host-guaranteed kv_blocks >= 2 so the Q slot outlives this read):
      // s_tail[t] = <q_row, k_tail_t> on CUDA cores, q from the swizzled smem tile, k_tail broadcast
      // from the item's tail slot.  Column half 0 owns the tail: its scores join that half's max and
      // sum of the LAST key block.
      float s_tail[kMaxTail];
#pragma unroll
      for (int t = 0; t < kMaxTail; ++t) s_tail[t] = -INFINITY;
      if constexpr (HD == 64) {
        if (p.k_tail > 0 && hf == 0) {
          mbar_wait(&q_full[item_i & 1], (item_i >> 1) & 1);
          const uint32_t q_row_addr =
              smem_u32(smem + Cfg::kOffQ + ((item_i & 1) * 2 + g) * Cfg::kTileBytes + r * 128);
          const uint32_t k_tail_addr = smem_u32(smem + Cfg::kOffTail + (item_i & 1) * Cfg::kTailSlotBytes);
#pragma unroll
          for (int t = 0; t < kMaxTail; ++t) {
            if (t >= p.k_tail) break;  // (unrolled over t only so that s_tail[] stays in registers)
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 1  // once per item: keep it small, it is always instruction-cache cold
            for (int c = 0; c < 8; ++c) {
              const float4 qv = lds128(q_row_addr + ((static_cast<uint32_t>(c) ^ sw) << 4));
              const float4 kv4 = lds128(k_tail_addr + t * 128 + c * 16);
              const float2 q0 = unpack_bf16x2(__float_as_uint(qv.x)), k0 = unpack_bf16x2(__float_as_uint(kv4.x));
              const float2 q1 = unpack_bf16x2(__float_as_uint(qv.y)), k1 = unpack_bf16x2(__float_as_uint(kv4.y));
              const float2 q2 = unpack_bf16x2(__float_as_uint(qv.z)), k2 = unpack_bf16x2(__float_as_uint(kv4.z));
              const float2 q3 = unpack_bf16x2(__float_as_uint(qv.w)), k3 = unpack_bf16x2(__float_as_uint(kv4.w));
              acc0 = fmaf(q0.x, k0.x, acc0); acc1 = fmaf(q0.y, k0.y, acc1);
              acc0 = fmaf(q1.x, k1.x, acc0); acc1 = fmaf(q1.y, k1.y, acc1);
              acc0 = fmaf(q2.x, k2.x, acc0); acc1 = fmaf(q2.y, k2.y, acc1);
              acc0 = fmaf(q3.x, k3.x, acc0); acc1 = fmaf(q3.y, k3.y, acc1);
            }
            s_tail[t] = acc0 + acc1;
          }
        }
      }
      float e_tail[kMaxTail];
#pragma unroll
      for (int t = 0; t < kMaxTail; ++t) e_tail[t] = 0.f;

      for (int j = 0; j < kvb; ++j, ++cnt) {
        const int valid_blk = min(128, p.Nk_main - j * 128);
        const int valid = max(0, min(64, valid_blk - hf * 64));  // valid columns of this half
        const int nchunks = (valid + 31) >> 5;
        const bool full_block = (valid == 64);
        mbar_wait(&s_full[g], cnt & 1);
        tc_fence_after();
        // half S row -> registers, then hand the TMEM buffer straight back to the MMA warp
        uint32_t sreg[64];
        if (warp_live) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
            if (c < nchunks) tmem_ld_32x32b_x32_p(t_s + c * 32, sreg + c * 32);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[g]);   // one arrival per warp: 256 per-thread arrivals serialise on the word

        // local max over this half's columns (8 independent chains), then the row max across halves
        float mx_loc = -INFINITY;
        if (warp_live) {
          if (full_block) {
            float mxh[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) mxh[a] = fmaxf(__uint_as_float(sreg[a]), __uint_as_float(sreg[a + 8]));
#pragma unroll
            for (int i = 16; i < 64; i += 16) {
#pragma unroll
              for (int a = 0; a < 8; ++a)
                mxh[a] = fmaxf(mxh[a], fmaxf(__uint_as_float(sreg[i + a]), __uint_as_float(sreg[i + a + 8])));
            }
            mx_loc = fmaxf(fmaxf(fmaxf(mxh[0], mxh[1]), fmaxf(mxh[2], mxh[3])),
                           fmaxf(fmaxf(mxh[4], mxh[5]), fmaxf(mxh[6], mxh[7])));
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (c < nchunks) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (c * 32 + i < valid) mx_loc = fmaxf(mx_loc, __uint_as_float(sreg[c * 32 + i]));
              }
            }
          }
          if (HD == 64 && j == kvb - 1) {
#pragma unroll
            for (int t = 0; t < kMaxTail; ++t) mx_loc = fmaxf(mx_loc, s_tail[t]);  // -inf when unused / half 1
          }
        }
        // exchange (double-buffered by block parity; one 64-thread named barrier per block and row quarter)
        float* xc = xchg + ((cnt & 1) * 4 + g * 2) * 128;
        xc[hf * 128 + r] = mx_loc;
        named_bar_sync(1 + g * 4 + quarter, 64);  // just the two warps that share these 32 rows
        const float mx_row = fmaxf(mx_loc, xc[(hf ^ 1) * 128 + r]);

        float alpha = 1.f;
        bool rescale = false;
        if (warp_live) {
          // identical in both halves: same row maxima, same warp quarter -> same lazy decision
          const float m_cand = fmaxf(m_run, mx_row * p.scale_log2);
          if (j == 0) {
            m_run = m_cand;  // nothing accumulated yet: P V(0) overwrites O
          } else if (__any_sync(0xffffffffu, m_cand > m_run + kLazyMaxLog2)) {
            alpha = fast_exp2(m_run - m_cand);
            l_run *= alpha;
            m_run = m_cand;
            rescale = true;
          }
          const float neg_m = -m_run;
          float sum = 0.f;
          if (full_block) sum = exp_row<true, POLY>(sreg, p.scale_log2, neg_m, valid, nchunks);
          else if (nchunks > 0) sum = exp_row<false, POLY>(sreg, p.scale_log2, neg_m, valid, nchunks);
          if (HD == 64 && j == kvb - 1) {
#pragma unroll
            for (int t = 0; t < kMaxTail; ++t) {
              e_tail[t] = fast_exp2(fmaf(s_tail[t], p.scale_log2, neg_m));  // exp2(-inf) = 0
              sum += e_tail[t];
            }
          }
          l_run += sum;
        }
        if (cnt > 0) {
          mbar_wait(&o_full[g], (cnt - 1) & 1);  // previous P V of this group retired: P_g free, O_g stable
          tc_fence_after();
        }
        if (rescale && hf == 0) {  // warp-uniform, rare; column half 0 rescales the shared accumulator
#pragma unroll 1
          for (int c = 0; c < HD / 16; ++c) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_o + c * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32b_x16(t_o + c * 16, v);
          }
        }
        if (warp_live) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
            if (c < nchunks) tmem_st_32x32b_x16_p(t_p + c * 16, sreg + c * 16);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
      }

      // row statistics for the epilogue warps (double-buffered by this group's item parity)
      if (ic > 0) mbar_wait(&o_free[g], (ic - 1) & 1);  // keeps stats_full at most one phase ahead
      {
        float* st = stats + ((ic & 1) * 2 + g) * (Cfg::kStatFields * 128);
        st[hf * 128 + r] = l_run;
        if (hf == 0) {
          st[2 * 128 + r] = m_run;
#pragma unroll
          for (int t = 0; t < kMaxTail; ++t) st[(3 + t) * 128 + r] = e_tail[t];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&stats_full[g]);
      ++ic;
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    int ic0 = 0, ic1 = 0;      // items finished per group (scalars: the group loop is rolled)
    int pcnt0 = 0, pcnt1 = 0;  // P V MMAs per group
    int item_i = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_i) {
      const int pair = item % p.q_pairs;
      const int bh = item / p.q_pairs;
      const int h = bh % p.H, b = bh / p.H;
      const int n_groups = groups_of(item);
      if (HD == 64 && p.k_tail > 0) mbar_wait(&q_full[item_i & 1], (item_i >> 1) & 1);
      const uint32_t v_tail_addr =
          smem_u32(smem + Cfg::kOffTail + (item_i & 1) * Cfg::kTailSlotBytes + kMaxTail * 128);
#pragma unroll 1
      for (int g = 0; g < n_groups; ++g) {
        const int qrow = (pair * 2 + g) * 128 + r;
        const int icg = g == 0 ? ic0 : ic1;
        const int pcg = (g == 0 ? pcnt0 : pcnt1) + kvb;
        if (g == 0) { pcnt0 = pcg; ++ic0; } else { pcnt1 = pcg; ++ic1; }
        const float* st = stats + ((icg & 1) * 2 + g) * (Cfg::kStatFields * 128);
        mbar_wait(&stats_full[g], icg & 1);
        mbar_wait(&o_full[g], (pcg - 1) & 1);  // the item's last P V retired
        tc_fence_after();
        const float l_row = st[0 * 128 + r] + st[1 * 128 + r];  // the two column halves
        const float inv_l = 1.f / l_row;
        const bool row_ok = qrow < p.Nq_main;
        const uint32_t t_o = tmem_base + lane_off + 256 + g * 64;
        __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.Nq + qrow) * p.ldo + h * HD;
#pragma unroll 1
        for (int c = 0; c < HD / 16; ++c) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_o + c * 16, v);
          tmem_ld_wait();
          if constexpr (HD == 64) {
#pragma unroll
            for (int t = 0; t < kMaxTail; ++t) {  // O += p_tail * v_tail (peeled keys)
              if (t >= p.k_tail) break;
              const float e = st[(3 + t) * 128 + r];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const float4 vv = lds128(v_tail_addr + t * 128 + c * 32 + u * 16);
                const float2 a0 = unpack_bf16x2(__float_as_uint(vv.x)), a1 = unpack_bf16x2(__float_as_uint(vv.y)),
                             a2 = unpack_bf16x2(__float_as_uint(vv.z)), a3 = unpack_bf16x2(__float_as_uint(vv.w));
                v[u * 8 + 0] = __float_as_uint(fmaf(e, a0.x, __uint_as_float(v[u * 8 + 0])));
                v[u * 8 + 1] = __float_as_uint(fmaf(e, a0.y, __uint_as_float(v[u * 8 + 1])));
                v[u * 8 + 2] = __float_as_uint(fmaf(e, a1.x, __uint_as_float(v[u * 8 + 2])));
                v[u * 8 + 3] = __float_as_uint(fmaf(e, a1.y, __uint_as_float(v[u * 8 + 3])));
                v[u * 8 + 4] = __float_as_uint(fmaf(e, a2.x, __uint_as_float(v[u * 8 + 4])));
                v[u * 8 + 5] = __float_as_uint(fmaf(e, a2.y, __uint_as_float(v[u * 8 + 5])));
                v[u * 8 + 6] = __float_as_uint(fmaf(e, a3.x, __uint_as_float(v[u * 8 + 6])));
                v[u * 8 + 7] = __float_as_uint(fmaf(e, a3.y, __uint_as_float(v[u * 8 + 7])));
              }
            }
          }
          if (row_ok) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              uint4 pk;
              pk.x = pack_bf16x2(__uint_as_float(v[t * 8 + 0]) * inv_l, __uint_as_float(v[t * 8 + 1]) * inv_l);
              pk.y = pack_bf16x2(__uint_as_float(v[t * 8 + 2]) * inv_l, __uint_as_float(v[t * 8 + 3]) * inv_l);
              pk.z = pack_bf16x2(__uint_as_float(v[t * 8 + 4]) * inv_l, __uint_as_float(v[t * 8 + 5]) * inv_l);
              pk.w = pack_bf16x2(__uint_as_float(v[t * 8 + 6]) * inv_l, __uint_as_float(v[t * 8 + 7]) * inv_l);
              *reinterpret_cast<uint4*>(orow + c * 16 + t * 8) = pk;
            }
          }
        }
        if (p.lse != nullptr && row_ok)
          p.lse[(static_cast<long long>(b) * p.H + h) * p.Nq + qrow] =
              (st[2 * 128 + r] + log2f(l_row)) * 0.6931471805599453f;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[g]);
      }
      if (HD == 64 && p.k_tail > 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&tail_free[item_i & 1]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// Peeled query rows [Nq_main, Nq) -- in MIRAGE the global token's query -- on CUDA cores.
// One CTA per (batch, head); K and V of that head are streamed once from L2/HBM (128-byte rows),
// so the kernel is bandwidth-bound: 2 * Nk * HD * 2 bytes per CTA.
//   phase 1  thread t scores keys t, t+128, ...: s = <q, k> * scale*log2(e)        (K rows: 8 x 16 B)
//   phase 2  block max / sum of exp2
//   phase 3  warp w accumulates keys w, w+4, ... for the dim pair (2*lane, 2*lane+1) (V rows coalesced)
// ---------------------------------------------------------------------------------------------
constexpr int kTailThreads = 128;
constexpr int kTailMaxKeys = 2048;

template <int HD>
__global__ void __launch_bounds__(kTailThreads)
attn_tail_rows_kernel(const AttnDev p) {
  static_assert(HD == 64, "tail rows are only peeled for head_dim 64");
  __shared__ float s_p[kTailMaxKeys];
  __shared__ float s_q[HD];
  __shared__ float s_red[8];
  __shared__ float s_o[4][HD];
  pdl_sync();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // CTAs walk the (batch, head) problems from the END: the tile kernel in front of this one finished with the last
  // batches, whose K / V rows are the ones still in the 126 MB L2
  const int bid = static_cast<int>(gridDim.x - 1 - blockIdx.x);
  const int h = bid % p.H, b = bid / p.H;
  const __nv_bfloat16* kbase = p.k + static_cast<long long>(b) * p.Nk * p.ldk + h * HD;
  const __nv_bfloat16* vbase = p.v + static_cast<long long>(b) * p.Nk * p.ldv + h * HD;

  for (int qrow = p.Nq_main; qrow < p.Nq; ++qrow) {
    __syncthreads();
    if (tid < HD)
      s_q[tid] = __bfloat162float(p.q[(static_cast<long long>(b) * p.Nq + qrow) * p.ldq + h * HD + tid]) *
                 p.scale_log2;
    __syncthreads();
    float mx = -INFINITY;
    for (int key = tid; key < p.Nk; key += kTailThreads) {
      const uint4* krow = reinterpret_cast<const uint4*>(kbase + static_cast<long long>(key) * p.ldk);
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int c = 0; c < HD / 8; ++c) {
        const uint4 kv4 = __ldg(krow + c);
        const float2 k0 = unpack_bf16x2(kv4.x), k1 = unpack_bf16x2(kv4.y), k2 = unpack_bf16x2(kv4.z),
                     k3 = unpack_bf16x2(kv4.w);
        acc0 = fmaf(s_q[c * 8 + 0], k0.x, acc0); acc1 = fmaf(s_q[c * 8 + 1], k0.y, acc1);
        acc0 = fmaf(s_q[c * 8 + 2], k1.x, acc0); acc1 = fmaf(s_q[c * 8 + 3], k1.y, acc1);
        acc0 = fmaf(s_q[c * 8 + 4], k2.x, acc0); acc1 = fmaf(s_q[c * 8 + 5], k2.y, acc1);
        acc0 = fmaf(s_q[c * 8 + 6], k3.x, acc0); acc1 = fmaf(s_q[c * 8 + 7], k3.y, acc1);
      }
      const float sc = acc0 + acc1;
      s_p[key] = sc;
      mx = fmaxf(mx, sc);
    }
    mx = warp_max(mx);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    float sum = 0.f;
    for (int key = tid; key < p.Nk; key += kTailThreads) {
      const float e = fast_exp2(s_p[key] - mx);
      s_p[key] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) s_red[4 + warp] = sum;
    __syncthreads();
    sum = s_red[4] + s_red[5] + s_red[6] + s_red[7];
    // P V: warp w takes keys w, w+4, ...; lane owns dims 2*lane, 2*lane+1
    float o0 = 0.f, o1 = 0.f;
#pragma unroll 8
    for (int key = warp; key < p.Nk; key += 4) {
      const uint32_t vv =
          __ldg(reinterpret_cast<const uint32_t*>(vbase + static_cast<long long>(key) * p.ldv) + lane);
      const float2 vf = unpack_bf16x2(vv);
      const float e = s_p[key];
      o0 = fmaf(e, vf.x, o0);
      o1 = fmaf(e, vf.y, o1);
    }
    s_o[warp][2 * lane] = o0;
    s_o[warp][2 * lane + 1] = o1;
    __syncthreads();
    if (tid < HD) {
      const float o = (s_o[0][tid] + s_o[1][tid] + s_o[2][tid] + s_o[3][tid]) / sum;
      p.out[(static_cast<long long>(b) * p.Nq + qrow) * p.ldo + h * HD + tid] = __float2bfloat16(o);
    }
    if (tid == 0 && p.lse != nullptr)
      p.lse[(static_cast<long long>(b) * p.H + h) * p.Nq + qrow] = (mx + log2f(sum)) * 0.6931471805599453f;
  }
}

// Same computation for FOUR consecutive heads per CTA: a token's K (or V) row of four heads is 512
// contiguous bytes, read by one warp with one 16-byte load per lane, instead of four CTAs each pulling
// a 128-byte piece of the same DRAM page at different times (ncu r01: 3.9 TB/s for the per-head kernel).
//   lane l: head l >> 3 of the group, dims [8 (l & 7), +8);  warp w: keys w, w+4, ...
template <int HD>
__global__ void __launch_bounds__(kTailThreads)
attn_tail_rows4_kernel(const AttnDev p) {
  static_assert(HD == 64, "tail rows are only peeled for head_dim 64");
  extern __shared__ float tsm[];
  pdl_sync();
  float* s_p = tsm;                        // [4][Nk_pad]
  const int nk_pad = (p.Nk + 3) & ~3;
  float* s_q = s_p + 4 * nk_pad;           // [4][64]   q * scale * log2(e)
  float* s_o = s_q + 4 * HD;               // [4 warps][4 heads * 64]
  float* s_ml = s_o + 4 * 4 * HD;          // [4] max, [4] sum
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hh = lane >> 3, c8 = (lane & 7) * 8;
  const int groups = p.H / 4;
  const int bid = static_cast<int>(gridDim.x - 1 - blockIdx.x);   // most recently used K / V first (see above)
  const int h0 = (bid % groups) * 4, b = bid / groups;
  const __nv_bfloat16* kbase = p.k + static_cast<long long>(b) * p.Nk * p.ldk + h0 * HD + lane * 8;
  const __nv_bfloat16* vbase = p.v + static_cast<long long>(b) * p.Nk * p.ldv + h0 * HD + lane * 8;

  for (int qrow = p.Nq_main; qrow < p.Nq; ++qrow) {
    __syncthreads();
    for (int i = tid; i < 4 * HD; i += kTailThreads)
      s_q[i] = __bfloat162float(p.q[(static_cast<long long>(b) * p.Nq + qrow) * p.ldq + h0 * HD + i]) * p.scale_log2;
    __syncthreads();
    float qf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qf[i] = s_q[hh * HD + c8 + i];
    // ---- scores
#pragma unroll 4
    for (int key = warp; key < p.Nk; key += 4) {
      const uint4 kv4 = __ldg(reinterpret_cast<const uint4*>(kbase + static_cast<long long>(key) * p.ldk));
      const float2 k0 = unpack_bf16x2(kv4.x), k1 = unpack_bf16x2(kv4.y), k2 = unpack_bf16x2(kv4.z),
                   k3 = unpack_bf16x2(kv4.w);
      float a = qf[0] * k0.x + qf[1] * k0.y + qf[2] * k1.x + qf[3] * k1.y + qf[4] * k2.x + qf[5] * k2.y +
                qf[6] * k3.x + qf[7] * k3.y;
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      if ((lane & 7) == 0) s_p[hh * nk_pad + key] = a;
    }
    __syncthreads();
    // ---- softmax of head `warp`
    {
      float* pr = s_p + warp * nk_pad;
      float mx = -INFINITY;
      for (int key = lane; key < p.Nk; key += 32) mx = fmaxf(mx, pr[key]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int key = lane; key < p.Nk; key += 32) {
        const float e = fast_exp2(pr[key] - mx);
        pr[key] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      if (lane == 0) {
        s_ml[warp] = mx;
        s_ml[4 + warp] = sum;
      }
    }
    __syncthreads();
    // ---- P V
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 4
    for (int key = warp; key < p.Nk; key += 4) {
      const uint4 vv = __ldg(reinterpret_cast<const uint4*>(vbase + static_cast<long long>(key) * p.ldv));
      const float e = s_p[hh * nk_pad + key];
      const float2 v0 = unpack_bf16x2(vv.x), v1 = unpack_bf16x2(vv.y), v2 = unpack_bf16x2(vv.z),
                   v3 = unpack_bf16x2(vv.w);
      acc[0] = fmaf(e, v0.x, acc[0]); acc[1] = fmaf(e, v0.y, acc[1]);
      acc[2] = fmaf(e, v1.x, acc[2]); acc[3] = fmaf(e, v1.y, acc[3]);
      acc[4] = fmaf(e, v2.x, acc[4]); acc[5] = fmaf(e, v2.y, acc[5]);
      acc[6] = fmaf(e, v3.x, acc[6]); acc[7] = fmaf(e, v3.y, acc[7]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s_o[warp * 4 * HD + lane * 8 + i] = acc[i];
    __syncthreads();
    for (int i = tid; i < 4 * HD; i += kTailThreads) {
      const int head = i / HD;
      const float o = (s_o[i] + s_o[4 * HD + i] + s_o[8 * HD + i] + s_o[12 * HD + i]) / s_ml[4 + head];
      p.out[(static_cast<long long>(b) * p.Nq + qrow) * p.ldo + h0 * HD + i] = __float2bfloat16(o);
    }
    if (tid < 4 && p.lse != nullptr)
      p.lse[(static_cast<long long>(b) * p.H + h0 + tid) * p.Nq + qrow] =
          (s_ml[tid] + log2f(s_ml[4 + tid])) * 0.6931471805599453f;
  }
}

template <int HD, int POLY>
static int launch_attn_fwd(const mb_attn_args* a, cudaStream_t stream) {
  using Cfg = AttnCfg<HD>;
  AttnDev p;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.lse = a->lse;
  p.q = reinterpret_cast<const __nv_bfloat16*>(a->q);
  p.k = reinterpret_cast<const __nv_bfloat16*>(a->k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(a->v);
  p.ldq = a->ldq;
  p.ldk = a->ldk;
  p.ldv = a->ldv;
  p.ldo = a->ldo;
  p.B = (int)a->batch;
  p.H = (int)a->heads;
  p.Nq = (int)a->nq;
  p.Nk = (int)a->nk;
  // peel a 1..kMaxTail remainder off the key / query ranges (see AttnDev)
  p.Nk_main = p.Nk;
  p.k_tail = 0;
  if (HD == 64 && p.Nk >= 256 && p.Nk % 128 >= 1 && p.Nk % 128 <= kMaxTail) {
    p.k_tail = p.Nk % 128;
    p.Nk_main = p.Nk - p.k_tail;
  }
  p.Nq_main = p.Nq;
  if (HD == 64 && p.Nq >= 128 && p.Nq % 128 >= 1 && p.Nq % 128 <= kMaxTail && p.Nk <= kTailMaxKeys)
    p.Nq_main = p.Nq - p.Nq % 128;
  p.q_tiles = (p.Nq_main + 127) / 128;
  p.q_pairs = (p.q_tiles + 1) / 2;
  p.kv_blocks = (p.Nk_main + 127) / 128;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.reverse = 0;
  {
    static int stagger = -1;
    if (stagger < 0) {
      const char* e = getenv("MB_ATTN_STAGGER");
      stagger = e ? atoi(e) : kDefaultStagger;
    }
    p.stagger = stagger;
  }
  bool main_done = false;
  if constexpr (HD == 64) {
    // one query tile x one key tile (the 99-token sequences of a pretraining step): attention_small.cu;
    // MB_ATTN_SMALL=0 disables it
    static int use_small = -1;
    if (use_small < 0) {
      const char* e = getenv("MB_ATTN_SMALL");
      use_small = e ? atoi(e) : 1;
    }
    if (use_small && p.Nq <= 128 && p.Nk <= 128) {
      const int rc = launch_attn_fwd_small(a, p, stream);
      if (rc <= 0) return rc;
    }
  }
  if constexpr (HD == 64) {
    // long sequences in whole quads of query tiles: the four-tile kernel (attention4.cu); MB_ATTN_V4=0 disables it
    static int use_v4 = -1;
    if (use_v4 < 0) {
      const char* e = getenv("MB_ATTN_V4");
      use_v4 = e ? atoi(e) : 1;
    }
    if (use_v4) {
      // a quarter of the exponentials on the FMA pipe (exp2_poly2) pays off in the four-tile kernel (0.57 -> 0.51 ms
      // at cfg 2) although it never did in the two-tile kernel below (MB_ATTN_POLY overrides both)
      static int poly4 = -1;
      if (poly4 < 0) {
        const char* e = getenv("MB_ATTN_POLY");
        poly4 = e ? atoi(e) : 4;
      }
      const int rc = launch_attn_fwd4(a, p, poly4, stream);
      if (rc < 0) return rc;
      main_done = (rc == 0);
    }
  }
  if (!main_done) {
  note_direction(+1);   // the two-tile kernel walks the problems upwards
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
  auto kern = attn_fwd_kernel<HD, POLY>;
  static PerDeviceOnce configured;
  if (configured.first()) {
    MB_CHECK_CUDA(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  }
  const long long items = (long long)p.B * p.H * p.q_pairs;
  MB_REQUIRE(items > 0 && items < (1ll << 31), "mb_attn_fwd: %lld work items out of range", items);
  const long long grid = items < sm_count() ? items : sm_count();  // persistent CTAs
  MB_CHECK_CUDA(launch_k(kern, dim3((unsigned)grid), dim3(kAttnThreads), Cfg::kSmemBytes, stream, tq, tk, tv, p));
  }
  if constexpr (HD == 64) {
    if (p.Nq_main < p.Nq) {
      // four heads per CTA read 512-byte row segments (better DRAM locality) but give 4x fewer CTAs: only
      // when that still fills the machine several times over
      if (p.H % 4 == 0 && p.ldk % 8 == 0 && p.ldv % 8 == 0 && (long long)p.B * (p.H / 4) >= 4ll * sm_count()) {
        const size_t tsmem = (size_t)(4 * ((p.Nk + 3) & ~3) + 4 * HD + 16 * HD + 8) * sizeof(float);
        static PerDeviceOnce tail_configured;
        if (tail_configured.first()) {
          MB_CHECK_CUDA(cudaFuncSetAttribute(attn_tail_rows4_kernel<HD>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024 + 16 * 1024));
        }
        MB_CHECK_CUDA(launch_row(attn_tail_rows4_kernel<HD>, dim3((unsigned)(p.B * (p.H / 4))), dim3(kTailThreads), tsmem,
                               stream, p));
      } else {
        MB_CHECK_CUDA(launch_row(attn_tail_rows_kernel<HD>, dim3((unsigned)(p.B * p.H)), dim3(kTailThreads), 0, stream, p));
      }
    }
    MB_CHECK_CUDA(cudaGetLastError());
  }
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
  // share of the exponentials evaluated on the FMA pipe (pairs out of every 16); tunable for
  // experiments through MB_ATTN_POLY = 0 | 4 | 6 | 8
  static int poly = -1;
  if (poly < 0) {
    const char* e = getenv("MB_ATTN_POLY");
    poly = e ? atoi(e) : kDefaultPoly;
  }
  if (a->head_dim == 64) {
    switch (poly) {
      case 0: return launch_attn_fwd<64, 0>(a, stream);
      case 6: return launch_attn_fwd<64, 6>(a, stream);
      case 8: return launch_attn_fwd<64, 8>(a, stream);
      default: return launch_attn_fwd<64, 4>(a, stream);
    }
  }
  return launch_attn_fwd<32, 0>(a, stream);
}
