// mirage_b200/csrc/attention4.cu
//
// Attention forward for the long MIRAGE sequences (N = 512 t + 1: encoder inference at 513 tokens, segmentation
// inputs at 1025 / 2049), head_dim 64:  O = softmax(Q K^T * scale) V  per (batch, head).
// Replaces F.scaled_dot_product_attention + transpose/reshape at mirage/utils.py:181-185.
//
// Why a second kernel.  attention.cu splits every score row over TWO threads (column halves) that exchange their
// row maxima through shared memory and a named barrier once per key block.  ncu (profiles/r01_ncu_attn_fwd_v6):
// that barrier is the single largest stall of the softmax warps (20 % of their samples), the two halves run in
// lockstep, and all four softmax warps of an SM sub-partition end up exponentiating at the same time and idling at
// the same time -- XU pipe 48 %, tensor pipe 23 %.  Here a row belongs to ONE thread and an SM works on FOUR
// 128-row query tiles at once, each with its own MMA issuer and barrier chain:
//
//   work item     (batch, head, four consecutive query tiles = 512 rows); keys walked in blocks of 64
//   warp 0        TMA producer: the item's 4 Q tiles (double-buffered across items) + K_j / V_j through 4-stage rings
//   warps 1-4     one MMA-issuing thread PER query tile g:  P_g V_j  then  S_g(j+1) = Q_g K_{j+1}^T
//   warps 8-23    softmax: warpgroup g owns tile g, thread r owns row r: 64 scores per block in registers
//                 (one tcgen05.ld round), thread-private running max / sum, lazy rescale, exponentials packed to
//                 bf16 and written back IN PLACE over the first half of S_g (tcgen05.st) -- no exchange, no named
//                 barrier, no shared memory on the per-block path
//   warps 24-27   epilogue: O_g / l -> bf16 -> global (+ LSE) after the item's last P V
//
// The four tiles are independent pipelines sharing one MUFU and one tensor core: while tile g waits for its
// P V -> S round trip (S and P alias, so S_g(j+1) cannot start before P_g V_j has consumed P_g), the other three
// exponentiate.  TMEM: S_g / P_g [64 g, 64 g + 64), O_g [256 + 64 g, 256 + 64 g + 64) -- all 512 columns.
//
// The global-token remainder is peeled exactly as in attention.cu: the +1 KEY is a rank-1 update on CUDA cores
// (score in the softmax thread, e * v_tail in the epilogue), the +1 QUERY row goes to attn_tail_rows4_kernel.
// (Computing that row here with two spare warps reading the staged K / V blocks -- lane = key for the scores, lane =
// output chunk for P V -- was implemented and measured twice in round 2: correct, but 0.63-0.65 ms vs 0.52 ms with
// the separate kernel at cfg 2.  The two extra arrivals on every K / V stage-release barrier tie the TMA ring of
// the four tiles to two latency-bound CUDA-core warps limited to 40 registers; round 1 saw the same in the
// two-tile kernel.)
#include "attention_common.cuh"

namespace mb200 {

constexpr int kA4GroupsDbg = 4;
#ifdef MB_ATTN4_TIMING
// debug build only (MB_NVCC_EXTRA=-DMB_ATTN4_TIMING): SM-clock timestamps of CTA 0, key blocks [kT0, kT0 + kTN)
constexpr int kT0 = 100, kTN = 6, kTE = 12;
__device__ long long g_a4_t[kTN][kA4GroupsDbg][4][kTE];
#define A4_T(blk, g, q, e)                                                                   \
  do {                                                                                       \
    if (blockIdx.x == 0 && (blk) >= kT0 && (blk) < kT0 + kTN) g_a4_t[(blk) - kT0][g][q][e] = clock64(); \
  } while (0)
#else
#define A4_T(blk, g, q, e) do { } while (0)
#endif

constexpr int kA4Threads = 896;     // 7 warpgroups: {TMA, MMA x3}, {MMA, -, -, -}, 4 x softmax, epilogue
constexpr int kA4Groups = 4;
constexpr int kA4KvStages = 4;
constexpr int kA4Block = 64;        // keys per block
constexpr int kA4MaxTail = 2;

struct A4Cfg {
  static constexpr int kQTile = 128 * 128;                     // 128 rows x 64 bf16
  static constexpr int kKvTile = kA4Block * 128;               // 64 keys x 64 bf16
  static constexpr int kOffQ = 0;                              // 2 item slots x 4 tiles
  static constexpr int kOffK = kOffQ + 2 * kA4Groups * kQTile;
  static constexpr int kOffV = kOffK + kA4KvStages * kKvTile;
  static constexpr int kOffTail = kOffV + kA4KvStages * kKvTile;   // 2 slots x {k rows, v rows}
  static constexpr int kTailSlot = 2 * kA4MaxTail * 128;
  static constexpr int kOffStats = kOffTail + 2 * kTailSlot;       // [parity][group][field][row] f32
  static constexpr int kStatFields = 2 + kA4MaxTail;               // l, m, e_tail[]
  static constexpr int kStatsBytes = 2 * kA4Groups * kStatFields * 128 * 4;
  static constexpr int kOffBar = kOffStats + kStatsBytes;
  static constexpr int kSmemBytes = kOffBar + 512;
};
static_assert(A4Cfg::kSmemBytes <= 227 * 1024, "attention4 shared memory exceeds the 227 KB per-CTA limit");

// One 64-column score row in registers -> packed bf16 exponentials in sreg[0..32), returns their sum and (e_max)
// their maximum.  Three passes in pinned program order as exp_row (attention_common.cuh): x = s * scale - m
// (FFMA2), e = 2^x in place (MUFU back to back), then sum (FADD2, four chains) + maximum (FMNMX, four chains)
// + bf16 packing.
//
// Tried and rejected (all measured on B200, cfg 2): pacing the MUFU passes of the four softmax warps
// of an SM sub-partition so that they do not bunch -- a token ring per sub-partition (1.14 ms with one token,
// 0.82-0.84 ms with two to four, vs 0.56 ms unpaced), a FIFO ticket lock making the pass exclusive (0.62 vs 0.60
// ms), and issuing the MMAs from thread 0 of the softmax warpgroup behind a named barrier instead of a dedicated
// issuer warp (0.64 ms).  scripts/micro/mufu_rate.cu shows why pacing cannot help: the MUFU is shared fairly and
// one warp alone already sustains 97 % of its rate, so the pass length is not the limiter.
template <int POLY>
__device__ __forceinline__ float exp_row_max(uint32_t (&sreg)[64], float scale_log2, float neg_m, float& e_max) {
#pragma unroll
  for (int i = 0; i < 64; i += 2) ffma2_v(sreg[i], sreg[i + 1], scale_log2, neg_m);
#pragma unroll
  for (int i = 0; i < 64; i += 2) {
    if (((i >> 1) * POLY) / 16 != (((i >> 1) + 1) * POLY) / 16) {
      float e0, e1;
      exp2_poly2(e0, e1, __uint_as_float(sreg[i]), __uint_as_float(sreg[i + 1]));
      sreg[i] = __float_as_uint(e0);
      sreg[i + 1] = __float_as_uint(e1);
    } else {
      ex2_v(sreg[i]);
      ex2_v(sreg[i + 1]);
    }
  }
  float sum[8], mx[4];
#pragma unroll
  for (int a = 0; a < 8; ++a) sum[a] = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) mx[a] = 0.f;
#pragma unroll
  for (int i = 0; i < 64; i += 2) {
    const int a = (i >> 1) & 3;
    fadd2_v(sum[2 * a], sum[2 * a + 1], sreg[i], sreg[i + 1]);
    mx[a] = fmaxf(mx[a], fmaxf(__uint_as_float(sreg[i]), __uint_as_float(sreg[i + 1])));
    sreg[i >> 1] = pack_bf16x2(__uint_as_float(sreg[i]), __uint_as_float(sreg[i + 1]));
  }
  e_max = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
  return ((sum[0] + sum[1]) + (sum[2] + sum[3])) + ((sum[4] + sum[5]) + (sum[6] + sum[7]));
}

template <int POLY>
__global__ void __launch_bounds__(kA4Threads, 1)
attn_fwd4_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const AttnDev p) {
  using Cfg = A4Cfg;
  constexpr int HD = 64;
  constexpr uint64_t kSw = kDescSwizzle128B;
  constexpr uint32_t kSbo = 8 * 128;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* q_full = bars;                      // 2   TMA tx (4 Q tiles + tail rows of the item)
  uint64_t* q_empty = bars + 2;                 // 2   4 commits: every issuer's last S of the item
  uint64_t* k_full = bars + 4;                  // 4
  uint64_t* k_empty = bars + 8;                 // 4   4 commits
  uint64_t* v_full = bars + 12;                 // 4
  uint64_t* v_empty = bars + 16;                // 4   4 commits
  uint64_t* s_full = bars + 20;                 // 4   MMA commit: S_g(j) complete (and P_g V(j-1) retired)
  uint64_t* p_full = bars + 24;                 // 4   4 softmax warps (one elected lane each): P_g(j) in TMEM
  uint64_t* o_full = bars + 28;                 // 4   MMA commit: the item's last P_g V retired
  uint64_t* o_free = bars + 32;                 // 4   4 epilogue warps: O_g of the item read out
  uint64_t* stats_full = bars + 36;             // 4   128 softmax threads: row statistics of the item
  uint64_t* tail_free = bars + 40;              // 2   4 epilogue warps: tail rows of the slot consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 42);
  float* stats = reinterpret_cast<float*>(smem + Cfg::kOffStats);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("mirage_b200: attention4 smem base not 1024-byte aligned\n");
    __trap();
  }

  const int kvb = p.kv_blocks;                  // 64-key blocks
  const int q_quads = p.q_tiles / kA4Groups;
  const int n_items = p.B * p.H * q_quads;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], kA4Groups);
      mbar_init(&tail_free[s], 4);
    }
    for (int s = 0; s < kA4KvStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], kA4Groups);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], kA4Groups);
    }
    for (int g = 0; g < kA4Groups; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 4);
      mbar_init(&o_full[g], 1);
      mbar_init(&o_free[g], 4);
      mbar_init(&stats_full[g], 128);
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

  if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && lane == 0) {
      // -------------------------------------------------------------- TMA producer
      int kv_count = 0, item_i = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_i) {
        const int ritem = p.reverse ? n_items - 1 - item : item;
        const int quad = ritem % q_quads;
        const int bh = ritem / q_quads;
        const int h = bh % p.H, b = bh / p.H;
        const int slot = item_i & 1;
        const uint32_t ph = (item_i >> 1) & 1;
        mbar_wait(&q_empty[slot], ph ^ 1);
        uint32_t bytes = kA4Groups * Cfg::kQTile;
        if (p.k_tail > 0) {
          mbar_wait(&tail_free[slot], ph ^ 1);
          bytes += 2 * p.k_tail * 128;
        }
        mbar_arrive_expect_tx(&q_full[slot], bytes);
        for (int g = 0; g < kA4Groups; ++g)
          tma_load_3d(smem + Cfg::kOffQ + (slot * kA4Groups + g) * Cfg::kQTile, &tm_q, &q_full[slot], h * HD,
                      (quad * kA4Groups + g) * 128, b);
        uint8_t* tslot = smem + Cfg::kOffTail + slot * Cfg::kTailSlot;
        for (int t = 0; t < p.k_tail; ++t) {
          const long long row = static_cast<long long>(b) * p.Nk + p.Nk_main + t;
          bulk_load(tslot + t * 128, p.k + row * p.ldk + h * HD, 128, &q_full[slot]);
          bulk_load(tslot + (kA4MaxTail + t) * 128, p.v + row * p.ldv + h * HD, 128, &q_full[slot]);
        }
        for (int j = 0; j < kvb; ++j, ++kv_count) {
          const int s = kv_count % kA4KvStages;
          const uint32_t kph = (kv_count / kA4KvStages) & 1;
          mbar_wait(&k_empty[s], kph ^ 1);
          mbar_arrive_expect_tx(&k_full[s], Cfg::kKvTile);
          tma_load_3d(smem + Cfg::kOffK + s * Cfg::kKvTile, &tm_k, &k_full[s], h * HD, j * kA4Block, b);
          mbar_wait(&v_empty[s], kph ^ 1);
          mbar_arrive_expect_tx(&v_full[s], Cfg::kKvTile);
          tma_load_3d(smem + Cfg::kOffV + s * Cfg::kKvTile, &tm_v, &v_full[s], h * HD, j * kA4Block, b);
        }
      }
      pdl_trigger();
    } else if (warp >= 1 && warp <= kA4Groups) {
      // -------------------------------------------------------------- MMA issuer of query tile g
      // Event order of one tile: S(0) | P(0) -> PV(0), S(1) | P(1) -> PV(1), S(2) | ...  S_g and P_g alias in
      // TMEM, and MMAs issued by one thread execute in order, so S(j+1) lands only after P V(j) has read P(j).
      // The next item's S(0) is issued right behind this item's last P V (its Q sits in the other slot).
      //
      // The softmax warps wait for the whole P(j) -> [P V(j), S(j+1)] round trip, so this path is kept short:
      //   * the WHOLE warp runs the loop (converged) and every address is derived from shuffled, i.e. provably
      //     warp-uniform, values: the MMA operands then live in uniform registers.  The first version ran
      //     under `lane == 0` with per-thread values, and ptxas wrapped every tcgen05.mma in an ELECT /
      //     R2UR.BROADCAST / BRA.U.ANY loop: ~100 clk per MMA, 1150-1400 clk from "P ready" to "S committed"
      //     (clock64 timeline, scripts/attn4_timeline.py);
      //   * everything that does not depend on P(j) -- V(j) / K(j+1) / next Q arrival, O_g drained -- is waited
      //     for BEFORE p_full, and all descriptors are formed before it too.
      const int g = __shfl_sync(0xffffffffu, warp, 0) - 1;
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const bool leader = elect_one();
      const uint32_t q_addr = smem_u32(smem + Cfg::kOffQ);
      const uint32_t k_addr = smem_u32(smem + Cfg::kOffK);
      const uint32_t v_addr = smem_u32(smem + Cfg::kOffV);
      const uint32_t t_s = tb + g * 64;
      const uint32_t t_o = tb + 256 + g * 64;
      constexpr uint32_t idesc_s = make_idesc(128, kA4Block, kFmtBF16, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(128, HD, kFmtBF16, 0, 1);
      int pc = 0;  // P V MMAs issued (phase of p_full)
      int oc = 0;  // items finished  (phase of o_free)

      // S(it, block kc): caller has made sure Q(it) is resident
      auto issue_s = [&](int it, int kc, bool last_of_item, bool wait_k) {
        const int s = kc % kA4KvStages;
        if (wait_k) {
          mbar_wait(&k_full[s], (kc / kA4KvStages) & 1);
          tc_fence_after();
        }
        const uint64_t qd = make_smem_desc(q_addr + ((it & 1) * kA4Groups + g) * Cfg::kQTile, 0, kSbo, kSw);
        const uint64_t kd = make_smem_desc(k_addr + s * Cfg::kKvTile, 0, kSbo, kSw);
        if (leader) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)   // +32 bytes per K step = +2 in the descriptor's 16-byte address field
            umma_f16_ss(t_s, qd + 2 * k, kd + 2 * k, idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&s_full[g]);
          umma_commit(&k_empty[s]);
          if (last_of_item) umma_commit(&q_empty[it & 1]);
        }
        __syncwarp();
      };

      int kv_base = 0, item_i = 0;
      if (blockIdx.x < n_items) {
        mbar_wait(&q_full[0], 0);
        issue_s(0, 0, false, true);
      }
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_i, kv_base += kvb) {
        const bool has_next = item + static_cast<int>(gridDim.x) < n_items;
        for (int j = 0; j < kvb; ++j) {
          const int kc = kv_base + j;
          const int sv = kc % kA4KvStages;
          const bool last = (j == kvb - 1);
          // --- not on the critical path: operands of the two MMA groups
          mbar_wait(&v_full[sv], (kc / kA4KvStages) & 1);
          if (!last || has_next) {
            if (last) mbar_wait(&q_full[(item_i + 1) & 1], ((item_i + 1) >> 1) & 1);
            mbar_wait(&k_full[(kc + 1) % kA4KvStages], ((kc + 1) / kA4KvStages) & 1);
          }
          if (j == 0 && oc > 0) mbar_wait(&o_free[g], (oc - 1) & 1);  // previous item's O read out
          const uint64_t vd = make_smem_desc(v_addr + sv * Cfg::kKvTile, 0, kSbo, kSw);
          A4_T(pc, g, 0, 2);
          // --- critical path: P(j) ready -> P V(j) -> S(j+1)
          mbar_wait(&p_full[g], pc & 1);
          tc_fence_after();
          A4_T(pc, g, 0, 1);
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < kA4Block / 16; ++kk)   // 16 keys = 2048 bytes = +128 in the address field
              umma_f16_ts(t_o, t_s + kk * 8, vd + 128 * kk, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&v_empty[sv]);
            if (last) umma_commit(&o_full[g]);
          }
          __syncwarp();
          A4_T(pc, g, 0, 3);
          ++pc;
          if (last) ++oc;
          if (!last) {
            issue_s(item_i, kc + 1, j + 1 == kvb - 1, false);
          } else if (has_next) {
            issue_s(item_i + 1, kc + 1, false, false);
          }
          A4_T(pc - 1, g, 0, 5);
        }
      }
    }
  } else if (warp < 24) {
    // ---------------------------------------------------------------- softmax: warpgroup g, thread = row
    // register pool of the CTA = 896 x 72 at launch = 64512 >= 128 x (3 x 40 + 4 x 96)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    const int g = (warp - 8) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                    // row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + lane_off + g * 64;
    const uint32_t t_o = tmem_base + lane_off + 256 + g * 64;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    int cnt = 0;   // key blocks processed (phase of s_full)
    int ic = 0;    // items processed     (stats slot, phase of o_free)
    int item_i = 0;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_i) {
      float m_run = -INFINITY;
      float l_run = 0.f;
      // peeled key tail: s_tail[t] = <q_row, k_tail_t> on CUDA cores (q from the swizzled smem tile)
      float s_tail[kA4MaxTail];
#pragma unroll
      for (int t = 0; t < kA4MaxTail; ++t) s_tail[t] = -INFINITY;
      if (p.k_tail > 0) {
        mbar_wait(&q_full[item_i & 1], (item_i >> 1) & 1);
        const uint32_t q_row_addr =
            smem_u32(smem + Cfg::kOffQ + ((item_i & 1) * kA4Groups + g) * Cfg::kQTile + r * 128);
        const uint32_t k_tail_addr = smem_u32(smem + Cfg::kOffTail + (item_i & 1) * Cfg::kTailSlot);
#pragma unroll
        for (int t = 0; t < kA4MaxTail; ++t) {
          if (t >= p.k_tail) break;
          float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 1
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
      float e_tail[kA4MaxTail];
#pragma unroll
      for (int t = 0; t < kA4MaxTail; ++t) e_tail[t] = 0.f;

      for (int j = 0; j < kvb; ++j, ++cnt) {
        mbar_wait(&s_full[g], cnt & 1);   // S_g(j) complete; P_g V(j-1) retired (O_g stable, P_g free)
        tc_fence_after();
        if (lane == 0) A4_T(cnt, g, quarter, 6);
        uint32_t sreg[64];
        tmem_ld_32x32b_x32_p(t_s, sreg);
        tmem_ld_32x32b_x32_p(t_s + 32, sreg + 32);
        tmem_ld_wait();
        if (lane == 0) A4_T(cnt, g, quarter, 7);

        // Running max, lazily: m_run is fixed by the first block and moves only when a later block exceeds it by
        // more than 2^8 (the final O / l does not depend on the reference point, and P <= 256 is well inside
        // bf16 / fp32 range).  Blocks j > 0 are therefore exponentiated SPECULATIVELY against the stale m_run
        // without looking for their maximum first -- that search (~400 clk incl. the vote, clock64 timeline) sat
        // on the per-block critical path; now the maximum is taken over the exponentials in the summation pass
        // (FMNMX on the ALU pipe, next to the FADD2 / pack instructions) and only a warp that finds one above
        // 2^8 (or inf) takes the slow path: reload S_g from TMEM (still intact: P_g is stored afterwards),
        // move m_run, rescale l and O_g, exponentiate again.
        if (j == 0) {
          float mxh[8];
#pragma unroll
          for (int a = 0; a < 8; ++a) mxh[a] = fmaxf(__uint_as_float(sreg[a]), __uint_as_float(sreg[a + 8]));
#pragma unroll
          for (int i = 16; i < 64; i += 16) {
#pragma unroll
            for (int a = 0; a < 8; ++a)
              mxh[a] = fmaxf(mxh[a], fmaxf(__uint_as_float(sreg[i + a]), __uint_as_float(sreg[i + a + 8])));
          }
          m_run = fmaxf(fmaxf(fmaxf(mxh[0], mxh[1]), fmaxf(mxh[2], mxh[3])),
                        fmaxf(fmaxf(mxh[4], mxh[5]), fmaxf(mxh[6], mxh[7]))) * p.scale_log2;
        }
        if (lane == 0) A4_T(cnt, g, quarter, 8);
        float e_max;
        float sum = exp_row_max<POLY>(sreg, p.scale_log2, -m_run, e_max);
        if (j == kvb - 1) {
#pragma unroll
          for (int t = 0; t < kA4MaxTail; ++t) {
            e_tail[t] = fast_exp2(fmaf(s_tail[t], p.scale_log2, -m_run));  // exp2(-inf) = 0
            e_max = fmaxf(e_max, e_tail[t]);
          }
        }
        if (__any_sync(0xffffffffu, !(e_max <= 256.f))) {
          // slow path (rare, warp-uniform)
          tmem_ld_32x32b_x32_p(t_s, sreg);
          tmem_ld_32x32b_x32_p(t_s + 32, sreg + 32);
          tmem_ld_wait();
          float mx_row = -INFINITY;
#pragma unroll
          for (int i = 0; i < 64; ++i) mx_row = fmaxf(mx_row, __uint_as_float(sreg[i]));
          if (j == kvb - 1) {
#pragma unroll
            for (int t = 0; t < kA4MaxTail; ++t) mx_row = fmaxf(mx_row, s_tail[t]);  // -inf when unused
          }
          const float m_new = fmaxf(m_run, mx_row * p.scale_log2);
          if (j > 0) {  // (j == 0 lands here only through the tail key of a two-block item)
            const float alpha = fast_exp2(m_run - m_new);
            l_run *= alpha;
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
          m_run = m_new;
          sum = exp_row_max<POLY>(sreg, p.scale_log2, -m_run, e_max);
          if (j == kvb - 1) {
#pragma unroll
            for (int t = 0; t < kA4MaxTail; ++t) e_tail[t] = fast_exp2(fmaf(s_tail[t], p.scale_log2, -m_run));
          }
        }
        if (lane == 0) A4_T(cnt, g, quarter, 9);
        if (j == kvb - 1) {
#pragma unroll
          for (int t = 0; t < kA4MaxTail; ++t) sum += e_tail[t];
        }
        l_run += sum;
        // P_g(j): 64 bf16 = 32 columns, written over the first half of S_g (this thread's own lane)
        tmem_st_32x32b_x16_p(t_s, sreg);
        tmem_st_32x32b_x16_p(t_s + 16, sreg + 16);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);   // one arrival per warp: 128 per-thread arrivals serialise on the word
        if (lane == 0) A4_T(cnt, g, quarter, 0);
      }

      // row statistics for the epilogue warps (double-buffered by item parity)
      if (ic > 0) mbar_wait(&o_free[g], (ic - 1) & 1);  // keeps stats_full at most one phase ahead
      {
        float* st = stats + ((ic & 1) * kA4Groups + g) * (Cfg::kStatFields * 128);
        st[r] = l_run;
        st[128 + r] = m_run;
#pragma unroll
        for (int t = 0; t < kA4MaxTail; ++t) st[(2 + t) * 128 + r] = e_tail[t];
      }
      mbar_arrive(&stats_full[g]);
      ++ic;
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    int item_i = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_i) {
      const int ritem = p.reverse ? n_items - 1 - item : item;
      const int quad = ritem % q_quads;
      const int bh = ritem / q_quads;
      const int h = bh % p.H, b = bh / p.H;
      if (p.k_tail > 0) mbar_wait(&q_full[item_i & 1], (item_i >> 1) & 1);
      const uint32_t v_tail_addr =
          smem_u32(smem + Cfg::kOffTail + (item_i & 1) * Cfg::kTailSlot + kA4MaxTail * 128);
#pragma unroll 1
      for (int g = 0; g < kA4Groups; ++g) {
        const int qrow = (quad * kA4Groups + g) * 128 + r;
        const float* st = stats + ((item_i & 1) * kA4Groups + g) * (Cfg::kStatFields * 128);
        mbar_wait(&stats_full[g], item_i & 1);
        mbar_wait(&o_full[g], item_i & 1);  // the item's last P V retired
        tc_fence_after();
        const float l_row = st[r];
        const float inv_l = 1.f / l_row;
        const uint32_t t_o = tmem_base + lane_off + 256 + g * 64;
        __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.Nq + qrow) * p.ldo + h * HD;
#pragma unroll 1
        for (int c = 0; c < HD / 16; ++c) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_o + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < kA4MaxTail; ++t) {  // O += p_tail * v_tail (peeled keys)
            if (t >= p.k_tail) break;
            const float e = st[(2 + t) * 128 + r];
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
        if (p.lse != nullptr)
          p.lse[(static_cast<long long>(b) * p.H + h) * p.Nq + qrow] =
              (st[128 + r] + log2f(l_row)) * 0.6931471805599453f;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[g]);
      }
      if (p.k_tail > 0) {
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

template <int POLY>
static int launch4(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnDev& p,
                   cudaStream_t stream) {
  auto kern = attn_fwd4_kernel<POLY>;
  static PerDeviceOnce configured;
  if (configured.first())
    MB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, A4Cfg::kSmemBytes));
  const long long items = static_cast<long long>(p.B) * p.H * (p.q_tiles / kA4Groups);
  MB_REQUIRE(items > 0 && items < (1ll << 31), "mb_attn_fwd: %lld work items out of range", items);
  const long long grid = items < sm_count() ? items : sm_count();
  MB_CHECK_CUDA(launch_k(kern, dim3(static_cast<unsigned>(grid)), dim3(kA4Threads), A4Cfg::kSmemBytes, stream, tq, tk,
                         tv, p));
  return 0;
}

// Fits: head_dim 64, whole quads of query tiles, whole 64-key blocks (>= 2), at most kA4MaxTail peeled keys.
int launch_attn_fwd4(const mb_attn_args* a, const AttnDev& p_in, int poly, cudaStream_t stream) {
  if (a->head_dim != 64) return 1;
  if (p_in.Nq_main <= 0 || p_in.Nq_main % (128 * kA4Groups) != 0) return 1;
  if (p_in.Nk_main < 2 * kA4Block || p_in.Nk_main % kA4Block != 0 || p_in.k_tail > kA4MaxTail) return 1;
  AttnDev p = p_in;
  p.reverse = take_direction() < 0 ? 1 : 0;
  p.q_tiles = p.Nq_main / 128;
  p.kv_blocks = p.Nk_main / kA4Block;
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
    uint32_t box[3] = {64u, (uint32_t)kA4Block, 1u};
    if (make_tensor_map(&tk, a->k, kTmaBF16, 3, dims, str, box, 128)) return -1;
    uint64_t strv[2] = {(uint64_t)a->ldv * 2, (uint64_t)a->nk * a->ldv * 2};
    if (make_tensor_map(&tv, a->v, kTmaBF16, 3, dims, strv, box, 128)) return -1;
  }
  switch (poly) {
    case 0: return launch4<0>(tq, tk, tv, p, stream);
    case 8: return launch4<8>(tq, tk, tv, p, stream);
    default: return launch4<4>(tq, tk, tv, p, stream);
  }
}

}  // namespace mb200

#ifdef MB_ATTN4_TIMING
extern "C" int mb_debug_attn4_timing(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, mb200::g_a4_t, sizeof(mb200::g_a4_t)) == cudaSuccess ? 0 : -1;
}
#endif
