// mirage_b200/csrc/attention.cu
//
// Fused flash-style attention forward on tcgen05/TMEM for sm_100a (no mask, no dropout):
//     O = softmax(Q K^T * scale) V        per (batch, head)
// Replaces F.scaled_dot_product_attention + the transpose/reshape copy at
//   mirage/utils.py:181-185 (Attention)  and  mirage/utils.py:216-220 (CrossAttention).
//
// Persistent CTAs (one per SM) walk work items = (batch, head, pair of 128-row query tiles); within an
// item the keys are walked in blocks of 128 (the last block is shortened to a multiple of 16 keys).
// The TMA producer and the MMA warp run ahead into the next item while the softmax warps finish the
// current one, so there is no per-item pipeline fill bubble.
//
//   warp 0      TMA producer: Q tiles once, then K_j / V_j through a 3-stage ring
//   warp 1      TMEM allocator + MMA issuer (one lane; warps 2-3 idle; the whole warpgroup gives its
//               registers to the softmax warps with setmaxnreg):
//                   S_g = Q_g K_j^T      (SS, both K-major,      accumulator S_g in TMEM)
//                   O_g += P_g V_j       (SS, P K-major from smem, V MN-major, accumulator O_g)
//   warps 4-7   softmax group 0 (one thread per query row, no shuffles, 224 registers)
//   warps 8-11  softmax group 1
//
// Each softmax thread pulls its whole S row (128 fp32) into registers with one round of tcgen05.ld and
// releases the TMEM buffer at once (s_free), so the MMA warp can already run S_g(j+1) while the row is
// being exponentiated; the MMA warp is an event-driven scheduler polling {s_free, p_full} of both
// groups, so neither group ever waits for the other.  Softmax runs in the log2 domain with a running
// max m and sum l per row (packed FFMA2/FADD2, masking only in the ragged last key block); O is
// rescaled in TMEM (tcgen05.ld / st) only when some row max in the warp actually moved.
//
// TMEM columns: S_0 [0,128)  S_1 [128,256)  O_0 [256,256+HD)  O_1 [320,320+HD).
#include "../../include/mirage_b200.h"
#include "common.cuh"

namespace mb200 {

struct AttnDev {
  __nv_bfloat16* out;
  float* lse;
  const __nv_bfloat16* q;  // raw operands: CUDA-core tail paths only
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  long long ldq, ldk, ldv, ldo;
  int B, H, Nq, Nk;
  // Rows / keys walked by the tensor-core tiles.  MIRAGE sequences are 128*t + 1 (the global token
  // is appended LAST, mirage/model.py:390-391): the one-row / one-key remainder would cost a whole
  // extra 128-row work item and an extra key block, so it is peeled off instead:
  //   k_tail keys   -> rank-1 update on CUDA cores inside the softmax threads of this kernel
  //   q_tail rows   -> attn_tail_rows_kernel (CUDA cores, one CTA per (batch, head))
  int Nq_main, Nk_main, k_tail;
  int q_tiles, q_pairs, kv_blocks;
  float scale_log2;
};

constexpr int kMaxTail = 4;  // largest remainder (mod 128) that is peeled off instead of padded

constexpr int kAttnThreads = 384;  // 3 warpgroups: {TMA, MMA, -, -}, softmax group 0, softmax group 1
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
  uint64_t* s_free = bars + 19;               // 2
  uint64_t* q_empty = bars + 21;              // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

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
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
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
      mbar_init(&s_free[g], 128);
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

  // Persistent CTA: every role walks the same static list of work items; all barrier phases are
  // tracked with running counters, so the producer / MMA warp run ahead into the next item while the
  // softmax warps are still finishing the current one.
  if (warp < 4) {
    // registers of this warpgroup go to the softmax warpgroups (a whole 128-wide S row lives there)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0 && lane == 0) {
      // -------------------------------------------------------------- TMA producer
      int kv_count = 0, item_i = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_i) {
        const int pair = item % p.q_pairs;
        const int bh = item / p.q_pairs;
        const int h = bh % p.H, b = bh / p.H;
        const int n_groups = (pair * 2 + 1 < p.q_tiles) ? 2 : 1;
        mbar_wait(q_empty, (item_i & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, n_groups * Cfg::kTileBytes);
        for (int g = 0; g < n_groups; ++g)
          tma_load_3d(smem + Cfg::kOffQ + g * Cfg::kTileBytes, &tm_q, q_full, h * HD,
                      (pair * 2 + g) * 128, b);
        for (int j = 0; j < kvb; ++j, ++kv_count) {
          const int s = kv_count % kKvStages;
          const uint32_t ph = (kv_count / kKvStages) & 1;
          mbar_wait(&k_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[s], Cfg::kTileBytes);
          tma_load_3d(smem + Cfg::kOffK + s * Cfg::kTileBytes, &tm_k, &k_full[s], h * HD, j * 128, b);
          mbar_wait(&v_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[s], Cfg::kTileBytes);
          tma_load_3d(smem + Cfg::kOffV + s * Cfg::kTileBytes, &tm_v, &v_full[s], h * HD, j * 128, b);
        }
      }
    } else if (warp == 1 && lane == 0) {
      // -------------------------------------------------------------- MMA issuer
      // One thread, blocking waits in the order the events occur when the two softmax groups
      // ping-pong half a block apart:
      //   S_0(0) S_1(0);  for j: { S_0(j+1), S_1(j+1)  (S buffer released: softmax holds the row in
      //   registers);  P_0 V(j), P_1 V(j)  (P tile written) }
      // (a polling scheduler over mbarrier.test_wait lost ~1000 clk per event to test_wait latency).
      const uint32_t q_addr = smem_u32(smem + Cfg::kOffQ);
      const uint32_t k_addr = smem_u32(smem + Cfg::kOffK);
      const uint32_t v_addr = smem_u32(smem + Cfg::kOffV);
      const uint32_t p_addr = smem_u32(smem + Cfg::kOffP);
      constexpr uint32_t idesc_pv = make_idesc(128, HD, kFmtBF16, 0, 1);
      int sc[2] = {0, 0};  // running count of S MMAs per group (phase of s_full / s_free)
      int pc[2] = {0, 0};  // running count of PV MMAs per group (phase of p_full / o_full)
      int kv_base = 0;     // running key-block counter at the start of the item (ring slot / phase)
      int item_i = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_i, kv_base += kvb) {
        const int pair = item % p.q_pairs;
        const int n_groups = (pair * 2 + 1 < p.q_tiles) ? 2 : 1;

        auto issue_s = [&](int g, int j) {
          const int kc = kv_base + j;
          if (sc[g] > 0) mbar_wait(&s_free[g], (sc[g] - 1) & 1);
          if (g == 0) mbar_wait(&k_full[kc % kKvStages], (kc / kKvStages) & 1);
          tc_fence_after();
          const int valid = min(128, p.Nk_main - j * 128);
          const uint32_t idesc_s = make_idesc(128, static_cast<uint32_t>((valid + 15) & ~15), kFmtBF16, 0, 0);
          const uint32_t qa = q_addr + g * Cfg::kTileBytes;
          const uint32_t ka = k_addr + (kc % kKvStages) * Cfg::kTileBytes;
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_f16_ss(tmem_base + g * 128, make_smem_desc(qa + k * 32, 0, kSbo, kSw),
                        make_smem_desc(ka + k * 32, 0, kSbo, kSw), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&s_full[g]);
          ++sc[g];
          if (g == n_groups - 1) {
            umma_commit(&k_empty[kc % kKvStages]);   // both groups' S(j) are issued
            if (j == kvb - 1) umma_commit(q_empty);  // last S of the item: the Q tiles may go
          }
        };
        auto issue_pv = [&](int g, int j) {
          const int kc = kv_base + j;
          mbar_wait(&p_full[g], pc[g] & 1);
          if (g == 0) mbar_wait(&v_full[kc % kKvStages], (kc / kKvStages) & 1);
          tc_fence_after();
          const int valid = min(128, p.Nk_main - j * 128);
          const int ksteps = (valid + 15) >> 4;
          const uint32_t pa = p_addr + g * Cfg::kPBytes;
          const uint32_t va = v_addr + (kc % kKvStages) * Cfg::kTileBytes;
          for (int kk = 0; kk < ksteps; ++kk)
            umma_f16_ss(tmem_base + 256 + g * 64,
                        make_smem_desc(pa + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024),
                        make_smem_desc(va + kk * 16 * Cfg::kRowBytes, 0, kSbo, kSw), idesc_pv,
                        (j > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&o_full[g]);
          ++pc[g];
          if (g == n_groups - 1) umma_commit(&v_empty[kc % kKvStages]);
        };

        mbar_wait(q_full, item_i & 1);
        for (int g = 0; g < n_groups; ++g) issue_s(g, 0);
        for (int j = 0; j < kvb; ++j) {
          if (j + 1 < kvb)
            for (int g = 0; g < n_groups; ++g) issue_s(g, j + 1);
          for (int g = 0; g < n_groups; ++g) issue_pv(g, j);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                    // row inside the tile == TMEM lane
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + g * 128;
    const uint32_t t_o = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + 256 + g * 64;
    uint8_t* p_row = smem + Cfg::kOffP + g * Cfg::kPBytes + r * 128;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    int cnt = 0;  // running count of key blocks this group has processed (barrier phases)
    int item_i = 0;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_i) {
      const int pair = item % p.q_pairs;
      const int bh = item / p.q_pairs;
      const int h = bh % p.H, b = bh / p.H;
      if (pair * 2 + g >= p.q_tiles) continue;            // this group has no tile in the item
      const int qrow = (pair * 2 + g) * 128 + r;          // query index inside this (b, h)
      const bool warp_live = (pair * 2 + g) * 128 + quarter * 32 < p.Nq_main;
      float m_run = -INFINITY;
      float l_run = 0.f;

      // Peeled key tail (HD == 64 only, host-guaranteed kv_blocks >= 2 so the Q tile outlives this
      // read): s_tail[t] = <q_row, k_tail_t> on CUDA cores, q from the swizzled smem tile, k broadcast
      // from global.  The scores join the LAST key block's max / sum; P_tail V_tail is added to O in
      // the epilogue.
      float s_tail[kMaxTail];
#pragma unroll
      for (int t = 0; t < kMaxTail; ++t) s_tail[t] = -INFINITY;
      if constexpr (HD == 64) {
        if (p.k_tail > 0) {
          mbar_wait(q_full, item_i & 1);
          const uint32_t q_row_addr = smem_u32(smem + Cfg::kOffQ + g * Cfg::kTileBytes + r * 128);
#pragma unroll
          for (int t = 0; t < kMaxTail; ++t) {
            if (t >= p.k_tail) break;
            const uint4* krow = reinterpret_cast<const uint4*>(
                p.k + (static_cast<long long>(b) * p.Nk + p.Nk_main + t) * p.ldk + h * HD);
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 qv = lds128(q_row_addr + ((static_cast<uint32_t>(c) ^ sw) << 4));
              const uint4 kv4 = __ldg(krow + c);
              const float2 q0 = unpack_bf16x2(__float_as_uint(qv.x)), k0 = unpack_bf16x2(kv4.x);
              const float2 q1 = unpack_bf16x2(__float_as_uint(qv.y)), k1 = unpack_bf16x2(kv4.y);
              const float2 q2 = unpack_bf16x2(__float_as_uint(qv.z)), k2 = unpack_bf16x2(kv4.z);
              const float2 q3 = unpack_bf16x2(__float_as_uint(qv.w)), k3 = unpack_bf16x2(kv4.w);
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
        const int valid = min(128, p.Nk_main - j * 128);
        const int nchunks = (valid + 31) >> 5;
        const bool full_block = (valid == 128);
        mbar_wait(&s_full[g], cnt & 1);
        tc_fence_after();
        // whole S row -> registers, then hand the TMEM buffer straight back to the MMA warp
        uint32_t sreg[128];
        if (warp_live) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < nchunks) tmem_ld_32x32b_x32_p(t_s + c * 32, sreg + c * 32);
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(&s_free[g]);

        float alpha = 1.f;
        float m_new = m_run;
        if (warp_live) {
          float mx0 = -INFINITY, mx1 = -INFINITY;
          if (full_block) {
#pragma unroll
            for (int i = 0; i < 128; i += 2) {
              mx0 = fmaxf(mx0, __uint_as_float(sreg[i]));
              mx1 = fmaxf(mx1, __uint_as_float(sreg[i + 1]));
            }
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (c < nchunks) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (c * 32 + i < valid) mx0 = fmaxf(mx0, __uint_as_float(sreg[c * 32 + i]));
              }
            }
          }
          if (HD == 64 && j == kvb - 1) {
#pragma unroll
            for (int t = 0; t < kMaxTail; ++t) mx0 = fmaxf(mx0, s_tail[t]);  // -inf when unused
          }
          m_new = fmaxf(m_run, fmaxf(mx0, mx1) * p.scale_log2);
          alpha = fast_exp2(m_run - m_new);
          // exponentiate in registers first (sreg[i/2] <- packed bf16 pair), so that the wait for the
          // previous block's P V (which still reads the P tile and writes O) comes as late as possible
          l_run *= alpha;
          float sum0 = 0.f, sum1 = 0.f;
          const float neg_m = -m_new;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c < nchunks) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const int idx = c * 32 + i;
                float x0, x1;
                ffma2(x0, x1, __uint_as_float(sreg[idx]), __uint_as_float(sreg[idx + 1]), p.scale_log2,
                      p.scale_log2, neg_m, neg_m);
                float e0 = fast_exp2(x0), e1 = fast_exp2(x1);
                if (!full_block) {
                  if (idx >= valid) e0 = 0.f;
                  if (idx + 1 >= valid) e1 = 0.f;
                }
                fadd2(sum0, sum1, sum0, sum1, e0, e1);
                sreg[idx >> 1] = pack_bf16x2(e0, e1);
              }
            }
          }
          if (HD == 64 && j == kvb - 1) {
#pragma unroll
            for (int t = 0; t < kMaxTail; ++t) {
              e_tail[t] = fast_exp2(fmaf(s_tail[t], p.scale_log2, neg_m));  // exp2(-inf) = 0
              sum0 += e_tail[t];
            }
          }
          l_run += sum0 + sum1;
          m_run = m_new;
        }
        if (j > 0) {
          mbar_wait(&o_full[g], (cnt - 1) & 1);  // P_g and O_g are free again
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
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c < nchunks) {
              uint8_t* dst = p_row + (c >> 1) * 16384;
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int w0 = c * 16 + t * 4;  // packed words of columns c*32 + t*8 .. +7
                const uint32_t c8 = static_cast<uint32_t>((c & 1) * 4 + t);
                *reinterpret_cast<uint4*>(dst + ((c8 ^ sw) << 4)) =
                    make_uint4(sreg[w0], sreg[w0 + 1], sreg[w0 + 2], sreg[w0 + 3]);
              }
            }
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(&p_full[g]);
      }

      // epilogue: O / l -> bf16 -> global, one 2*HD-byte row per thread
      mbar_wait(&o_full[g], (cnt - 1) & 1);
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
          if constexpr (HD == 64) {
#pragma unroll
            for (int t = 0; t < kMaxTail; ++t) {  // O += p_tail * v_tail (peeled keys)
              if (t >= p.k_tail) break;
              const uint4* vrow = reinterpret_cast<const uint4*>(
                  p.v + (static_cast<long long>(b) * p.Nk + p.Nk_main + t) * p.ldv + h * HD + c * 32);
              const float e = e_tail[t];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint4 vv = __ldg(vrow + u);
                const float2 a0 = unpack_bf16x2(vv.x), a1 = unpack_bf16x2(vv.y),
                             a2 = unpack_bf16x2(vv.z), a3 = unpack_bf16x2(vv.w);
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
          if (qrow < p.Nq_main) {
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
        if (p.lse != nullptr && qrow < p.Nq_main)
          p.lse[(static_cast<long long>(b) * p.H + h) * p.Nq + qrow] =
              (m_run + log2f(l_run)) * 0.6931471805599453f;
      }
      tc_fence_before();  // O_g reads retire before the next item's first P V overwrites it
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
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x % p.H, b = blockIdx.x / p.H;
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
  auto kern = attn_fwd_kernel<HD>;
  static bool configured = false;
  if (!configured) {
    MB_CHECK_CUDA(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const long long items = (long long)p.B * p.H * p.q_pairs;
  MB_REQUIRE(items > 0 && items < (1ll << 31), "mb_attn_fwd: %lld work items out of range", items);
  const long long grid = items < sm_count() ? items : sm_count();  // persistent CTAs
  kern<<<(unsigned)grid, kAttnThreads, Cfg::kSmemBytes, stream>>>(tq, tk, tv, p);
  MB_CHECK_CUDA(cudaGetLastError());
  if constexpr (HD == 64) {
    if (p.Nq_main < p.Nq)
      attn_tail_rows_kernel<HD><<<(unsigned)(p.B * p.H), kTailThreads, 0, stream>>>(p);
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
  if (a->head_dim == 64) return launch_attn_fwd<64>(a, stream);
  return launch_attn_fwd<32>(a, stream);
}
