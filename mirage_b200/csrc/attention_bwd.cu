// mirage_b200/csrc/attention_bwd.cu
//
// Attention backward on tcgen05/TMEM for sm_100a (no mask, no dropout), given the forward's
// log-sum-exp:   P = exp(S*scale - LSE),  dV = P^T dO,  dP = dO V^T,  dS = P o (dP - D) * scale,
//                dQ = dS K,  dK = dS^T Q,   D_i = rowsum(dO_i o O_i).
// Backward of F.scaled_dot_product_attention at mirage/utils.py:181-185 / :216-220.
//
// Persistent CTAs (one per SM) walk work items = (batch, head); inside an item the outer loop runs over
// key blocks j (128 keys), the inner loop over query tiles i (128 rows).  dK_j / dV_j accumulate in
// TMEM across the inner loop; dQ_i is produced per (j, i) and (when there is more than one key block)
// accumulated in an fp32 workspace by the thread that owns the row -- no atomics anywhere.
// The (j, i) iterations of ALL items form one software pipeline (the first version ran one CTA per
// item with every stage serialised: 228 us for the 4096 (b, h) problems of cfg 4, 15k clk per item):
//
//   warp 0       TMA producer: K_j, V_j per key block (2 slots); Q_i, dO_i, O_i per iteration (2 slots)
//   warp 1       MMA issuer (one lane):  MMA1(t+1) = {S, dP}  is issued as soon as the compute warps have
//                copied S/dP(t) to registers, i.e. it overlaps the exponentials of iteration t and MMA2(t)
//   warps 4-7    compute, key columns [0, 64)   of the tile: one thread per query row holds 64 S + 64 dP
//   warps 8-11   compute, key columns [64, 128)  values, writes bf16 P and dS into 128B-swizzled smem
//   warps 12-15  drain: dV / dK (end of a key block) and dQ (every iteration) TMEM -> global; the
//                accumulators are handed back to the MMA warp right after the TMEM reads, before the stores
//
// The same swizzled P / dS tile is consumed twice: as a K-major A operand (dQ = dS K) and, through
// the MN-major descriptor, as the transposed A operand (dV = P^T dO, dK = dS^T Q) -- no transposes.
//
// TMEM columns: S [0,128)  dP [128,256)  dV [256,+HD)  dK [320,+HD)  dQ [384,+HD).
#include "../../include/mirage_b200.h"
#include "common.cuh"

namespace mb200 {

struct AttnBwdDev {
  const __nv_bfloat16* o;
  const __nv_bfloat16* d_o;
  const float* lse;
  __nv_bfloat16* dq;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  float* dq_acc;  // fp32 [B*Nq, H*HD] when kv_blocks > 1
  long long ldo, lddo, lddq, lddk, lddv;
  int B, H, Nq, Nk;
  int q_tiles, kv_blocks;
  float scale, scale_log2;
};

constexpr int kAttnBwdThreads = 512;

template <int HD>
struct AttnBwdCfg {
  static constexpr int kRowBytes = HD * 2;
  static constexpr int kTileBytes = 128 * kRowBytes;
  static constexpr int kPBytes = 128 * 128 * 2;
  static constexpr int kOffK = 0;                    // 2 slots each
  static constexpr int kOffV = kOffK + 2 * kTileBytes;
  static constexpr int kOffQ = kOffV + 2 * kTileBytes;
  static constexpr int kOffDO = kOffQ + 2 * kTileBytes;
  static constexpr int kOffO = kOffDO + 2 * kTileBytes;   // forward output tile: D = rowsum(dO o O) from smem
  static constexpr int kOffP = kOffO + 2 * kTileBytes;
  static constexpr int kOffDS = kOffP + kPBytes;
  static constexpr int kOffBar = kOffDS + kPBytes;
  static constexpr int kSmemBytes = kOffBar + 256;
  static constexpr int kSwizzle = (HD == 64) ? 128 : 64;
};

template <int HD>
__global__ void __launch_bounds__(kAttnBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_o, const AttnBwdDev p) {
  using Cfg = AttnBwdCfg<HD>;
  constexpr uint64_t kSw = (HD == 64) ? kDescSwizzle128B : kDescSwizzle64B;
  constexpr uint32_t kSbo = 8 * Cfg::kRowBytes;
  constexpr uint32_t kRowStep16 = 16 * Cfg::kRowBytes;  // 16 rows of an operand tile

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* kv_full = bars + 0;    // 2  TMA tx
  uint64_t* kv_empty = bars + 2;   // 2  MMA commit after the key block's last MMA2
  uint64_t* qdo_full = bars + 4;   // 2  TMA tx
  uint64_t* qdo_empty = bars + 6;  // 2  MMA commit after MMA2(t)
  uint64_t* sdp_full = bars + 8;   // MMA commit: S, dP of iteration t complete
  uint64_t* sdp_free = bars + 9;   // 8 compute warps: S, dP copied to registers
  uint64_t* pds_full = bars + 10;  // 8 compute warps: P, dS tiles written
  uint64_t* mma2_done = bars + 11; // MMA commit: dV, dK, dQ of iteration t complete, P/dS tiles free
  uint64_t* acc_free = bars + 12;  // 4 drain warps: accumulators read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("mirage_b200: attention-bwd smem base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_o);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
      mbar_init(&qdo_full[s], 1);
      mbar_init(&qdo_empty[s], 1);
    }
    mbar_init(sdp_full, 1);
    // one arrival per WARP (an elected lane behind __syncwarp): 256 + 256 + 128 per-thread arrivals per iteration
    // serialise on three shared-memory words that all sit on the kernel's critical chain
    mbar_init(sdp_free, 8);
    mbar_init(pds_full, 8);
    mbar_init(mma2_done, 1);
    mbar_init(acc_free, 4);
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
  const int kvb = p.kv_blocks, qt = p.q_tiles;
  const int n_items = p.B * p.H;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    // Producer and issuer run as whole, converged warps on warp-uniform values and elect one lane only around the
    // TMA / MMA / commit instructions: under a `lane == 0` guard ptxas wraps every UTCHMMA / UTMALDG in an ELECT /
    // R2UR.BROADCAST / BRA.U.ANY loop (~100 clk of serial issue each, see attention4.cu) -- with up to 32 MMAs per
    // (key block, query tile) iteration against ~1300 clk of tensor work, issue WAS this kernel's critical path.
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer
      const bool leader = elect_one();
      int t = 0, jb = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int h = item % p.H, b = item / p.H;
        for (int j = 0; j < kvb; ++j, ++jb) {
          const int ks = jb & 1;
          mbar_wait(&kv_empty[ks], ((jb >> 1) & 1) ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(&kv_full[ks], 2 * Cfg::kTileBytes);
            tma_load_3d(smem + Cfg::kOffK + ks * Cfg::kTileBytes, &tm_k, &kv_full[ks], h * HD, j * 128, b);
            tma_load_3d(smem + Cfg::kOffV + ks * Cfg::kTileBytes, &tm_v, &kv_full[ks], h * HD, j * 128, b);
          }
          __syncwarp();
          for (int i = 0; i < qt; ++i, ++t) {
            const int qs = t & 1;
            mbar_wait(&qdo_empty[qs], ((t >> 1) & 1) ^ 1);
            if (leader) {
              mbar_arrive_expect_tx(&qdo_full[qs], 3 * Cfg::kTileBytes);
              tma_load_3d(smem + Cfg::kOffQ + qs * Cfg::kTileBytes, &tm_q, &qdo_full[qs], h * HD, i * 128, b);
              tma_load_3d(smem + Cfg::kOffDO + qs * Cfg::kTileBytes, &tm_do, &qdo_full[qs], h * HD, i * 128, b);
              tma_load_3d(smem + Cfg::kOffO + qs * Cfg::kTileBytes, &tm_o, &qdo_full[qs], h * HD, i * 128, b);
            }
            __syncwarp();
          }
        }
      }
      pdl_trigger();
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer
      const bool leader = elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t k_addr0 = smem_u32(smem + Cfg::kOffK);
      const uint32_t v_addr0 = smem_u32(smem + Cfg::kOffV);
      const uint32_t q_addr0 = smem_u32(smem + Cfg::kOffQ);
      const uint32_t do_addr0 = smem_u32(smem + Cfg::kOffDO);
      const uint32_t p_addr = smem_u32(smem + Cfg::kOffP);
      const uint32_t ds_addr = smem_u32(smem + Cfg::kOffDS);
      constexpr uint32_t idesc_dkv = make_idesc(128, HD, kFmtBF16, 1, 1);
      constexpr uint32_t idesc_dq = make_idesc(128, HD, kFmtBF16, 0, 1);
      const int iters_per_item = kvb * qt;
      const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int T = my_items * iters_per_item;

      // MMA1(t): S = Q_i K_j^T ; dP = dO_i V_j^T   (K-major operands, N = ncols)
      auto mma1 = [&](int t) {
        const int within = t % iters_per_item;
        const int j = within / qt, i = within - j * qt;
        const int jb = (t / iters_per_item) * kvb + j;
        const int valid = min(128, p.Nk - j * 128);
        const uint32_t idesc_s = make_idesc(128, static_cast<uint32_t>((valid + 15) & ~15), kFmtBF16, 0, 0);
        if (t > 0) mbar_wait(sdp_free, (t - 1) & 1);
        if (i == 0) mbar_wait(&kv_full[jb & 1], (jb >> 1) & 1);
        mbar_wait(&qdo_full[t & 1], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t q_addr = q_addr0 + (t & 1) * Cfg::kTileBytes, do_addr = do_addr0 + (t & 1) * Cfg::kTileBytes;
        const uint32_t k_addr = k_addr0 + (jb & 1) * Cfg::kTileBytes, v_addr = v_addr0 + (jb & 1) * Cfg::kTileBytes;
        // K step k = +32 bytes = +2 in the descriptor's 16-byte address field
        const uint64_t qd = make_smem_desc(q_addr, 0, kSbo, kSw), kd = make_smem_desc(k_addr, 0, kSbo, kSw);
        const uint64_t dod = make_smem_desc(do_addr, 0, kSbo, kSw), vd = make_smem_desc(v_addr, 0, kSbo, kSw);
        if (leader) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_f16_ss(tmem_u + 0, qd + 2 * k, kd + 2 * k, idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_f16_ss(tmem_u + 128, dod + 2 * k, vd + 2 * k, idesc_s, k > 0 ? 1u : 0u);
          umma_commit(sdp_full);
        }
        __syncwarp();
      };

      if (T > 0) mma1(0);
      for (int t = 0; t < T; ++t) {
        if (t + 1 < T) mma1(t + 1);
        const int within = t % iters_per_item;
        const int j = within / qt, i = within - j * qt;
        const int jb = (t / iters_per_item) * kvb + j;
        const int valid = min(128, p.Nk - j * 128);
        const uint32_t q_addr = q_addr0 + (t & 1) * Cfg::kTileBytes, do_addr = do_addr0 + (t & 1) * Cfg::kTileBytes;
        const uint32_t k_addr = k_addr0 + (jb & 1) * Cfg::kTileBytes;
        mbar_wait(pds_full, t & 1);
        if (t > 0) mbar_wait(acc_free, (t - 1) & 1);
        tc_fence_after();
        // dV_j += P^T dO_i ; dK_j += dS^T Q_i   (A = P / dS read MN-major: M = keys, K = query rows)
        // (descriptor address field in 16-byte units: +2048 bytes = +128, +kRowStep16 bytes = +kRowStep16 / 16)
        const uint64_t pd = make_smem_desc(p_addr, 16384, 1024), dsd = make_smem_desc(ds_addr, 16384, 1024);
        const uint64_t dsq = make_smem_desc(ds_addr, 0, 1024);
        const uint64_t dod = make_smem_desc(do_addr, 0, kSbo, kSw), qd = make_smem_desc(q_addr, 0, kSbo, kSw);
        const uint64_t kd = make_smem_desc(k_addr, 0, kSbo, kSw);
        const int ksteps = (valid + 15) >> 4;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_f16_ss(tmem_u + 256, pd + 128 * kk, dod + kk * (kRowStep16 >> 4), idesc_dkv,
                        (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_f16_ss(tmem_u + 320, dsd + 128 * kk, qd + kk * (kRowStep16 >> 4), idesc_dkv,
                        (i > 0 || kk > 0) ? 1u : 0u);
          // dQ_i(j) = dS K_j                      (A = dS K-major: M = query rows, K = keys)
          for (int kk = 0; kk < ksteps; ++kk)
            umma_f16_ss(tmem_u + 384, dsq + (kk >> 2) * 1024 + (kk & 3) * 2, kd + kk * (kRowStep16 >> 4), idesc_dq,
                        kk > 0 ? 1u : 0u);
          umma_commit(mma2_done);
          umma_commit(&qdo_empty[t & 1]);
          if (i == qt - 1) umma_commit(&kv_empty[jb & 1]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 12) {
    // ---------------------------------------------------------------- compute warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
    const int half = (warp - 4) >> 2;  // which 64 key columns of the tile
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    uint8_t* p_row = smem + Cfg::kOffP + half * 16384 + r * 128;
    uint8_t* ds_row = smem + Cfg::kOffDS + half * 16384 + r * 128;
    const float log2e = 1.4426950408889634f;
    int t = 0;
    // LSE of this thread's row, fetched one iteration ahead (a 4-byte global load per iteration whose
    // latency would otherwise sit in front of the exponentials)
    auto load_lse = [&](int tt) -> float {
      const int per_item = kvb * qt;
      const int it_item = blockIdx.x + (tt / per_item) * gridDim.x;
      if (it_item >= n_items) return 0.f;
      const int qr = ((tt % per_item) % qt) * 128 + r;
      if (qr >= p.Nq) return 0.f;
      return p.lse[(static_cast<long long>(it_item / p.H) * p.H + it_item % p.H) * p.Nq + qr];
    };
    float lse_next = load_lse(0);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int h = item % p.H, b = item / p.H;
      (void)h;
      (void)b;
      for (int j = 0; j < kvb; ++j) {
        const int valid = min(128, p.Nk - j * 128) - half * 64;  // valid columns of this half (may be <= 0)
        for (int i = 0; i < qt; ++i, ++t) {
          const int qrow = i * 128 + r;
          (void)qrow;
          // per-row statistics: LSE (global, 4 bytes) and D = <dO, O> from the TMA-staged tiles in smem (the
          // first version read both 128-byte rows from global memory here: that load latency was the
          // largest stall of the compute warps, ncu r01 attn_bwd v2).  Rows past Nq: the tiles are
          // zero-filled and any finite LSE keeps their P / dS harmless.
          const float lse2 = lse_next * log2e;
          float dsum = 0.f;
          lse_next = load_lse(t + 1);
          mbar_wait(&qdo_full[t & 1], (t >> 1) & 1);
          {
            const uint32_t o_row = smem_u32(smem + Cfg::kOffO + (t & 1) * Cfg::kTileBytes) + r * Cfg::kRowBytes;
            const uint32_t d_row = smem_u32(smem + Cfg::kOffDO + (t & 1) * Cfg::kTileBytes) + r * Cfg::kRowBytes;
            const uint32_t xr = (HD == 64) ? static_cast<uint32_t>(r & 7) : static_cast<uint32_t>((r >> 1) & 3);
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
              const uint32_t off = (static_cast<uint32_t>(c) ^ xr) << 4;
              const float4 a = lds128(o_row + off);
              const float4 g = lds128(d_row + off);
              const float2 a0 = unpack_bf16x2(__float_as_uint(a.x)), a1 = unpack_bf16x2(__float_as_uint(a.y)),
                           a2 = unpack_bf16x2(__float_as_uint(a.z)), a3 = unpack_bf16x2(__float_as_uint(a.w));
              const float2 g0 = unpack_bf16x2(__float_as_uint(g.x)), g1 = unpack_bf16x2(__float_as_uint(g.y)),
                           g2 = unpack_bf16x2(__float_as_uint(g.z)), g3 = unpack_bf16x2(__float_as_uint(g.w));
              dsum += a0.x * g0.x + a0.y * g0.y + a1.x * g1.x + a1.y * g1.y + a2.x * g2.x + a2.y * g2.y +
                      a3.x * g3.x + a3.y * g3.y;
            }
          }
          mbar_wait(sdp_full, t & 1);
          tc_fence_after();
          uint32_t vs[64], vd[64];
          tmem_ld_32x32b_x32_p(tmem_base + lane_off + half * 64, vs);
          tmem_ld_32x32b_x32_p(tmem_base + lane_off + half * 64 + 32, vs + 32);
          tmem_ld_32x32b_x32_p(tmem_base + lane_off + 128 + half * 64, vd);
          tmem_ld_32x32b_x32_p(tmem_base + lane_off + 128 + half * 64 + 32, vd + 32);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(sdp_free);  // MMA1(t+1) may overwrite S / dP now
          // P and dS in place: vs[e/2] <- packed P pair, vd[e/2] <- packed dS pair
          const float neg_lse = -lse2;
          const float ds_scale = p.scale;
#pragma unroll
          for (int e = 0; e < 64; e += 2) {
            const float x0 = fmaf(__uint_as_float(vs[e]), p.scale_log2, neg_lse);
            const float x1 = fmaf(__uint_as_float(vs[e + 1]), p.scale_log2, neg_lse);
            float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
            float d0 = p0 * (__uint_as_float(vd[e]) - dsum) * ds_scale;
            float d1 = p1 * (__uint_as_float(vd[e + 1]) - dsum) * ds_scale;
            if (valid < 64) {  // ragged key block: masked keys get P = dS = 0 (their S / dP columns were
                               // never written by the MMA and may hold anything, NaN included)
              if (e >= valid) p0 = d0 = 0.f;
              if (e + 1 >= valid) p1 = d1 = 0.f;
            }
            vs[e >> 1] = pack_bf16x2(p0, p1);
            vd[e >> 1] = pack_bf16x2(d0, d1);
          }
          if (t > 0) mbar_wait(mma2_done, (t - 1) & 1);  // MMA2(t-1) has finished reading the P / dS tiles
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            const uint32_t off = (static_cast<uint32_t>(c8) ^ sw) << 4;
            *reinterpret_cast<uint4*>(p_row + off) = make_uint4(vs[c8 * 4], vs[c8 * 4 + 1], vs[c8 * 4 + 2], vs[c8 * 4 + 3]);
            *reinterpret_cast<uint4*>(ds_row + off) = make_uint4(vd[c8 * 4], vd[c8 * 4 + 1], vd[c8 * 4 + 2], vd[c8 * 4 + 3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(pds_full);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- drain warps
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    int t = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int h = item % p.H, b = item / p.H;
      for (int j = 0; j < kvb; ++j) {
        const int krow = j * 128 + r;
        for (int i = 0; i < qt; ++i, ++t) {
          const int qrow = i * 128 + r;
          const bool row_ok = qrow < p.Nq;
          mbar_wait(mma2_done, t & 1);
          tc_fence_after();
          if (i == qt - 1) {
            // dV_j, dK_j complete: TMEM -> bf16 -> global (stores are fire-and-forget)
#pragma unroll
            for (int which = 0; which < 2; ++which) {
              __nv_bfloat16* base = which == 0 ? p.dv : p.dk;
              const long long ld = which == 0 ? p.lddv : p.lddk;
              __nv_bfloat16* dst = base + (static_cast<long long>(b) * p.Nk + krow) * ld + h * HD;
              uint32_t v[HD];  // one TMEM round trip per accumulator (the MMA warp waits for this drain)
#pragma unroll
              for (int c = 0; c < HD / 32; ++c)
                tmem_ld_32x32b_x32_p(tmem_base + lane_off + 256 + which * 64 + c * 32, v + c * 32);
              tmem_ld_wait();
              if (krow < p.Nk) {
#pragma unroll
                for (int q4 = 0; q4 < HD / 8; ++q4) {
                  uint4 pk;
                  pk.x = pack_bf16x2(__uint_as_float(v[q4 * 8 + 0]), __uint_as_float(v[q4 * 8 + 1]));
                  pk.y = pack_bf16x2(__uint_as_float(v[q4 * 8 + 2]), __uint_as_float(v[q4 * 8 + 3]));
                  pk.z = pack_bf16x2(__uint_as_float(v[q4 * 8 + 4]), __uint_as_float(v[q4 * 8 + 5]));
                  pk.w = pack_bf16x2(__uint_as_float(v[q4 * 8 + 6]), __uint_as_float(v[q4 * 8 + 7]));
                  *reinterpret_cast<uint4*>(dst + q4 * 8) = pk;
                }
              }
            }
          }
          // dQ_i(j): all TMEM reads first, then release the accumulators, then the global traffic
          uint32_t v[HD];
#pragma unroll
          for (int c = 0; c < HD / 32; ++c) tmem_ld_32x32b_x32_p(tmem_base + lane_off + 384 + c * 32, v + c * 32);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_free);
          if (row_ok) {
            const long long grow = static_cast<long long>(b) * p.Nq + qrow;
            if (kvb > 1) {
              float* acc = p.dq_acc + grow * (static_cast<long long>(p.H) * HD) + h * HD;
              if (j > 0) {
#pragma unroll
                for (int e = 0; e < HD; e += 4) {
                  const float4 a = *reinterpret_cast<const float4*>(acc + e);
                  v[e + 0] = __float_as_uint(__uint_as_float(v[e + 0]) + a.x);
                  v[e + 1] = __float_as_uint(__uint_as_float(v[e + 1]) + a.y);
                  v[e + 2] = __float_as_uint(__uint_as_float(v[e + 2]) + a.z);
                  v[e + 3] = __float_as_uint(__uint_as_float(v[e + 3]) + a.w);
                }
              }
              if (j < kvb - 1) {
#pragma unroll
                for (int e = 0; e < HD; e += 4)
                  *reinterpret_cast<float4*>(acc + e) =
                      make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]),
                                  __uint_as_float(v[e + 3]));
              }
            }
            if (j == kvb - 1) {
              __nv_bfloat16* dst = p.dq + grow * p.lddq + h * HD;
#pragma unroll
              for (int q4 = 0; q4 < HD / 8; ++q4) {
                uint4 pk;
                pk.x = pack_bf16x2(__uint_as_float(v[q4 * 8 + 0]), __uint_as_float(v[q4 * 8 + 1]));
                pk.y = pack_bf16x2(__uint_as_float(v[q4 * 8 + 2]), __uint_as_float(v[q4 * 8 + 3]));
                pk.z = pack_bf16x2(__uint_as_float(v[q4 * 8 + 4]), __uint_as_float(v[q4 * 8 + 5]));
                pk.w = pack_bf16x2(__uint_as_float(v[q4 * 8 + 6]), __uint_as_float(v[q4 * 8 + 7]));
                *reinterpret_cast<uint4*>(dst + q4 * 8) = pk;
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int HD>
static int launch_attn_bwd(const mb_attn_bwd_args* a, cudaStream_t stream) {
  using Cfg = AttnBwdCfg<HD>;
  CUtensorMap tq, tk, tv, tdo, to;
  uint32_t box[3] = {(uint32_t)HD, 128u, 1u};
  {
    uint64_t dims[3] = {(uint64_t)a->heads * HD, (uint64_t)a->nq, (uint64_t)a->batch};
    uint64_t s1[2] = {(uint64_t)a->ldq * 2, (uint64_t)a->nq * a->ldq * 2};
    if (make_tensor_map(&tq, a->q, kTmaBF16, 3, dims, s1, box, Cfg::kSwizzle)) return -1;
    uint64_t s2[2] = {(uint64_t)a->lddo * 2, (uint64_t)a->nq * a->lddo * 2};
    if (make_tensor_map(&tdo, a->d_out, kTmaBF16, 3, dims, s2, box, Cfg::kSwizzle)) return -1;
    uint64_t s3[2] = {(uint64_t)a->ldo * 2, (uint64_t)a->nq * a->ldo * 2};
    if (make_tensor_map(&to, a->out, kTmaBF16, 3, dims, s3, box, Cfg::kSwizzle)) return -1;
  }
  {
    uint64_t dims[3] = {(uint64_t)a->heads * HD, (uint64_t)a->nk, (uint64_t)a->batch};
    uint64_t s1[2] = {(uint64_t)a->ldk * 2, (uint64_t)a->nk * a->ldk * 2};
    if (make_tensor_map(&tk, a->k, kTmaBF16, 3, dims, s1, box, Cfg::kSwizzle)) return -1;
    uint64_t s2[2] = {(uint64_t)a->ldv * 2, (uint64_t)a->nk * a->ldv * 2};
    if (make_tensor_map(&tv, a->v, kTmaBF16, 3, dims, s2, box, Cfg::kSwizzle)) return -1;
  }
  AttnBwdDev p;
  p.o = reinterpret_cast<const __nv_bfloat16*>(a->out);
  p.d_o = reinterpret_cast<const __nv_bfloat16*>(a->d_out);
  p.lse = a->lse;
  p.dq = reinterpret_cast<__nv_bfloat16*>(a->dq);
  p.dk = reinterpret_cast<__nv_bfloat16*>(a->dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(a->dv);
  p.dq_acc = reinterpret_cast<float*>(a->workspace);
  p.ldo = a->ldo;
  p.lddo = a->lddo;
  p.lddq = a->lddq;
  p.lddk = a->lddk;
  p.lddv = a->lddv;
  p.B = (int)a->batch;
  p.H = (int)a->heads;
  p.Nq = (int)a->nq;
  p.Nk = (int)a->nk;
  p.q_tiles = (p.Nq + 127) / 128;
  p.kv_blocks = (p.Nk + 127) / 128;
  p.scale = a->scale;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  MB_REQUIRE(p.kv_blocks == 1 || a->workspace != nullptr,
             "mb_attn_bwd: nk > 128 needs a workspace of mb_attn_bwd_workspace() bytes");
  auto kern = attn_bwd_kernel<HD>;
  static PerDeviceOnce configured;
  if (configured.first()) {
    MB_CHECK_CUDA(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  }
  note_direction(+1);   // problems are walked upwards (take_direction(), common.cuh)
  const long long items = (long long)p.B * p.H;
  const long long grid = items < sm_count() ? items : sm_count();  // persistent CTAs
  MB_CHECK_CUDA(launch_k(kern, dim3((unsigned)grid), dim3(kAttnBwdThreads), Cfg::kSmemBytes, stream, tq, tk, tv, tdo,
                         to, p));
  return 0;
}

}  // namespace mb200

using namespace mb200;

extern "C" int64_t mb_attn_bwd_workspace(int64_t batch, int64_t heads, int64_t nq, int64_t nk,
                                         int32_t head_dim) {
  if (nk <= 128) return 0;
  return batch * nq * heads * head_dim * (int64_t)sizeof(float);
}

extern "C" int mb_attn_bwd(const mb_attn_bwd_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(a != nullptr, "mb_attn_bwd: null args");
  MB_REQUIRE(a->q && a->k && a->v && a->out && a->d_out && a->lse && a->dq && a->dk && a->dv,
             "mb_attn_bwd: null tensor pointer");
  MB_REQUIRE(a->batch > 0 && a->heads > 0 && a->nq > 0 && a->nk > 0, "mb_attn_bwd: empty problem");
  MB_REQUIRE(a->head_dim == 64 || a->head_dim == 32, "mb_attn_bwd: head_dim %d unsupported (32|64)",
             a->head_dim);
  MB_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0 &&
                 a->lddo % 8 == 0 && a->lddq % 8 == 0 && a->lddk % 8 == 0 && a->lddv % 8 == 0,
             "mb_attn_bwd: leading dimensions must be multiples of 8 elements");
  if (a->head_dim == 64) return launch_attn_bwd<64>(a, stream);
  return launch_attn_bwd<32>(a, stream);
}
