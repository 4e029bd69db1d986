// mirage_b200/csrc/gemm_impl.cuh
//
// Persistent, warp-specialised tcgen05 GEMMs for sm_100a:   D[M,N] = epilogue(A * B^T).
//
//   warp 0      TMA producer   (one lane): global -> 128B-swizzled smem ring, mbarrier tx-count
//   warp 1      MMA issuer     (one lane): tcgen05.mma kind::f16 / kind::tf32, 32 bytes of K per instr
//   warp 2      TMEM allocator (2 accumulator stages of BN fp32 columns)
//   warp 3      idle
//   warps 4-11  epilogue: tcgen05.ld -> per-warp swizzled smem slab -> coalesced global accesses
//
// Two kernels share everything but the tile / pairing logic:
//   gemm_kernel       one CTA per 128 x BN tile            (cta_group::1; small or ragged problems)
//   gemm_pair_kernel  one CTA pair per 256 x BN tile       (cta_group::2; the hot path)
// In the pair kernel each CTA stages its own 128 rows of A and only HALF of the B tile; the pair's
// tensor cores read both halves, so every byte of B is fetched from L2 once per 256 output rows.
// That cuts L2->SM operand traffic from 48 KB to 32 KB per SM per 64-deep K step -- the 1-CTA kernel
// is bound by exactly that traffic at large M (ncu: lts at its ~6.3 KB/clk ceiling, tensor pipe 38 %).
//
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the main loop of tile
// i+1.  Tiles are walked n-fastest so the CTAs resident at any moment share a few A row-panels in L2.
//
// The epilogue is a template (EPI_*): each variant is straight-line code with no dead branches -- the
// first version took every option at run time and spent most of its time in instruction-cache misses
// (ncu: stall_no_inst dominant), which capped the K=1024 GEMMs at half the main-loop rate.
#pragma once

#include "../../include/mirage_b200.h"
#include "common.cuh"

namespace mb200 {

struct GemmDev {
  void* out;
  const float* bias;
  const float* residual;
  const __nv_bfloat16* aux_in;
  __nv_bfloat16* aux_out;
  float* colsum_out;  // EPI_DGELU: column sums of the output accumulate here (bias gradient)
  // folded LayerNorm (see mb_gemm_args): producer side / consumer side
  __nv_bfloat16* twin_out;
  long long ld_twin;
  float* row_stats;
  const float* ln_stats;
  const float* ln_c1;
  float ln_eps;
  int ln_parts;      // row statistics: [M][1 + ln_parts][2] = total, then one partial sum per BN / 2 output columns
  int M, N, K;
  long long ldc, ld_res, ld_aux;
  int res_period;
  int out_f32;
  int epilogue;
  int m_tiles, n_tiles, k_blocks, k_splits, kb_per_split;
  int total_tiles;
  int m_reverse;     // row tiles are walked downwards (see take_direction(), common.cuh)
  // MB_A_PATCH32
  int rows_per_img;  // tokens per image
  int grid_w;        // patches per image row
  // output row map
  int orow_period;
  long long orow_stride, orow_offset;
  // MB_EPI_UNPATCH: token-major [B*gh*gw, C*ph*pw] -> image [B, C, gh*ph, gw*pw]
  int up_c, up_ph, up_pw, up_gh, up_gw;
};

// epilogue variants (compile-time)
enum : int {
  EPI_BF16 = 0,     // (+bias) -> bf16                                   qkv / q / kv / dgrad
  EPI_GELU = 1,     // (+bias) -> [aux_out = bf16 pre-activation] -> GELU -> bf16          fc1
  EPI_RES = 2,      // (+bias) + fp32 residual[row] -> fp32              proj / fc2 (in place allowed)
  EPI_DGELU = 3,    // * GELU'(aux_in) -> bf16                           dgrad through the GELU
  EPI_F32 = 4,      // (+bias) -> fp32, plain store or atomic add        wgrad (split-K), fp32 outputs
  EPI_GENERIC = 5,  // everything at run time (any combination; slow: integer divisions per chunk)
  EPI_MAP = 6,      // (+bias) (+fp32 residual, optionally periodic = pos-emb rows) -> fp32 through an
                    // output map (row map or un-patchify), index math hoisted out of the element loops:
                    // patch / semseg embedding into the token buffer, out_proj into image layout
  EPI_RES_LN = 7,   // EPI_RES + bf16 twin of the output + per-row {sum, sum of squares} (atomics)   proj / fc2
  EPI_BF16_LN = 8,  // LayerNorm folded in: rstd * acc - rstd * mean * c1[n] + c2[n] -> bf16               qkv
  EPI_GELU_LN = 9,  // same, then GELU -> bf16                                                           fc1
  EPI_COUNT = 10
};

constexpr bool epi_has_res(int mode) { return mode == EPI_RES || mode == EPI_RES_LN; }
constexpr bool epi_ln_in(int mode) { return mode == EPI_BF16_LN || mode == EPI_GELU_LN; }

constexpr int kGemmThreads = 384;
constexpr int kBM = 128;

template <int BN>
struct GemmCfg {  // one CTA per tile
  static constexpr int kABytes = kBM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kBarBytes = 256;
  static constexpr int kEpiBytes = 8 * 4096;  // one 32-row x 128-byte staging slab per epilogue warp
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + kEpiBytes + 1024;
};

template <int BN>
struct PairCfg {  // CTA pair per tile: each CTA holds BN/2 rows of B
  static constexpr int kABytes = kBM * 128;
  static constexpr int kBBytes = (BN / 2) * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 6 : 8;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kBarBytes = 256;
  static constexpr int kEpiBytes = 8 * 4096;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + kEpiBytes + 1024;
};

// ---------------------------------------------------------------------------------------------
// Epilogue of one warp: NSLABS slabs of 32 accumulator columns for the warp's 32 rows.
// Two phases per slab, private to the warp (only __syncwarp needed):
//   A  tcgen05.ld: thread = accumulator row; raw fp32 -> 128B-swizzled smem slab (conflict-free)
//   B  read the slab back "row-contiguous" (8 lanes per 128-byte row, 4 rows per instruction) and do
//      the element-wise math on 4 consecutive columns with fully coalesced 8/16-byte global accesses.
// Operands of phase B that come from global memory (bias, residual, GELU pre-activations) are
// fetched one slab ahead so their latency hides behind the previous slab.
// ---------------------------------------------------------------------------------------------
template <int MODE>
struct EpiRegs {
  float4 bias;
  float4 c1;  // folded LayerNorm: column sums of gamma * W
  float4 res[epi_has_res(MODE) ? 8 : 1];
  uint2 aux[MODE == EPI_DGELU ? 8 : 1];
};

template <int MODE>
__device__ __forceinline__ void epi_prefetch(const GemmDev& p, EpiRegs<MODE>& r, int row_base, int col,
                                             bool first_split, int sub) {
  r.bias = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool col_ok = col < p.N;
  if (p.bias != nullptr && first_split && col_ok)
    r.bias = __ldg(reinterpret_cast<const float4*>(p.bias + col));
  if constexpr (epi_ln_in(MODE)) {
    r.c1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col_ok) r.c1 = __ldg(reinterpret_cast<const float4*>(p.ln_c1 + col));
  }
  if constexpr (epi_has_res(MODE)) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = row_base + it * 4 + sub;
      r.res[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < p.M && col_ok && first_split)
        r.res[it] = *reinterpret_cast<const float4*>(p.residual + (long long)row * p.ld_res + col);
    }
  }
  if constexpr (MODE == EPI_DGELU) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = row_base + it * 4 + sub;
      r.aux[it] = make_uint2(0u, 0u);
      if (row < p.M && col_ok)
        r.aux[it] = __ldg(reinterpret_cast<const uint2*>(p.aux_in + (long long)row * p.ld_aux + col));
    }
  }
}

// run-time everything (EPI_GENERIC): one 4-column chunk
static __device__ __noinline__ void epi_generic_chunk(const GemmDev& p, float4 a4, int row, int col,
                                               bool first_split) {
  float f0 = a4.x, f1 = a4.y, f2 = a4.z, f3 = a4.w;
  if (p.bias != nullptr && first_split) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    f0 += b.x; f1 += b.y; f2 += b.z; f3 += b.w;
  }
  if (p.epilogue & MB_EPI_GELU) {
    if (p.aux_out) {
      uint2 pk;
      pk.x = pack_bf16x2(f0, f1);
      pk.y = pack_bf16x2(f2, f3);
      *reinterpret_cast<uint2*>(p.aux_out + (long long)row * p.ld_aux + col) = pk;
    }
    gelu_fast2(f0, f1); gelu_fast2(f2, f3);
  }
  if (p.epilogue & MB_EPI_DGELU) {
    const uint2 pk = __ldg(reinterpret_cast<const uint2*>(p.aux_in + (long long)row * p.ld_aux + col));
    float2 h0 = unpack_bf16x2(pk.x), h1 = unpack_bf16x2(pk.y);
    gelu_fast_grad2(h0.x, h0.y);
    gelu_fast_grad2(h1.x, h1.y);
    f0 *= h0.x; f1 *= h0.y; f2 *= h1.x; f3 *= h1.y;
  }
  if (p.residual != nullptr && first_split) {
    const long long rr = p.res_period > 0 ? row % p.res_period : row;
    const float4 r = *reinterpret_cast<const float4*>(p.residual + rr * p.ld_res + col);
    f0 += r.x; f1 += r.y; f2 += r.z; f3 += r.w;
  }
  long long oidx;
  if (p.epilogue & MB_EPI_UNPATCH) {
    const int per_img = p.up_gh * p.up_gw;
    const int bi = row / per_img, t = row - bi * per_img;
    const int nh = t / p.up_gw, nw = t - nh * p.up_gw;
    const int pp = p.up_ph * p.up_pw;
    const int ch = col / pp, rr2 = col - ch * pp;
    const int py = rr2 / p.up_pw, px = rr2 - py * p.up_pw;
    const long long W = (long long)p.up_gw * p.up_pw, H = (long long)p.up_gh * p.up_ph;
    oidx = (((long long)bi * p.up_c + ch) * H + (long long)nh * p.up_ph + py) * W +
           (long long)nw * p.up_pw + px;
  } else {
    const long long orow =
        p.orow_period > 0
            ? (long long)(row / p.orow_period) * p.orow_stride + row % p.orow_period + p.orow_offset
            : (long long)row;
    oidx = orow * p.ldc + col;
  }
  if (p.epilogue & MB_EPI_ATOMIC) {
    red_add_v4(reinterpret_cast<float*>(p.out) + oidx, f0, f1, f2, f3);  // oidx % 4 == 0 (ldc % 8, col % 4)
  } else if (p.out_f32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + oidx) = make_float4(f0, f1, f2, f3);
  } else {
    uint2 pk;
    pk.x = pack_bf16x2(f0, f1);
    pk.y = pack_bf16x2(f2, f3);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + oidx) = pk;
  }
}

template <int MODE, int NSLABS>
__device__ __forceinline__ void epilogue_warp(const GemmDev& p, uint32_t stage_addr, uint32_t taddr0,
                                              int row_base, int n_base, int split, int lane,
                                              uint64_t* full_bar, uint32_t full_parity) {
  const int sub = lane >> 3;  // phase B: row within a group of 4
  const int j4 = lane & 7;    // phase B: 16-byte chunk (4 columns) within the slab row
  const bool first_split = (split == 0);
  const uint32_t wr_addr = stage_addr + lane * 128;
  const uint32_t wr_sw = static_cast<uint32_t>(lane & 7);

  EpiRegs<MODE> cur, nxt;
  if constexpr (MODE != EPI_GENERIC && MODE != EPI_MAP)
    epi_prefetch<MODE>(p, cur, row_base, n_base + j4 * 4, first_split, sub);
  // EPI_MAP: row part of the output index and residual row of this thread's 8 rows, once per tile
  [[maybe_unused]] long long map_out[MODE == EPI_MAP ? 8 : 1];
  [[maybe_unused]] long long map_res[MODE == EPI_MAP ? 8 : 1];
  [[maybe_unused]] const bool unpatch = (p.epilogue & MB_EPI_UNPATCH) != 0;
  if constexpr (MODE == EPI_MAP) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = row_base + it * 4 + sub;
      map_res[it] = p.res_period > 0 ? row % p.res_period : row;
      if (unpatch) {
        const int per_img = p.up_gh * p.up_gw;
        const int bi = row / per_img, t = row - bi * per_img;
        const int nh = t / p.up_gw, nw = t - nh * p.up_gw;
        map_out[it] = ((static_cast<long long>(bi) * p.up_c) * (p.up_gh * p.up_ph) + static_cast<long long>(nh) * p.up_ph) *
                          (static_cast<long long>(p.up_gw) * p.up_pw) +
                      static_cast<long long>(nw) * p.up_pw;
      } else {
        const long long orow =
            p.orow_period > 0
                ? static_cast<long long>(row / p.orow_period) * p.orow_stride + row % p.orow_period + p.orow_offset
                : static_cast<long long>(row);
        map_out[it] = orow * p.ldc;
      }
    }
  }

  // folded LayerNorm, consumer side: rstd and rstd * mean of this thread's 8 rows (statistics over K columns)
  [[maybe_unused]] float ln_rstd[epi_ln_in(MODE) ? 8 : 1], ln_mr[epi_ln_in(MODE) ? 8 : 1];
  if constexpr (epi_ln_in(MODE)) {
    const float inv_k = 1.f / static_cast<float>(p.K);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = row_base + it * 4 + sub;
      float2 st = make_float2(0.f, 1.f);
      // slot 0 of the row = the producer's partial sums added in slot order by ln_stats_reduce_kernel (gemm.cu)
      if (row < p.M)
        st = __ldg(reinterpret_cast<const float2*>(p.ln_stats) + static_cast<long long>(row) * (p.ln_parts + 1));
      const float mean = st.x * inv_k;
      const float var = fmaxf(st.y * inv_k - mean * mean, 0.f);
      ln_rstd[it] = rsqrtf(var + p.ln_eps);
      ln_mr[it] = ln_rstd[it] * mean;
    }
  }
  // producer side: row sums of this thread's 4-column chunks over all slabs of the tile
  [[maybe_unused]] float rs1[MODE == EPI_RES_LN ? 8 : 1], rs2[MODE == EPI_RES_LN ? 8 : 1];
  if constexpr (MODE == EPI_RES_LN) {
#pragma unroll
    for (int it = 0; it < 8; ++it) rs1[it] = rs2[it] = 0.f;
  }

  mbar_wait(full_bar, full_parity);  // accumulator of this tile is complete
  tc_fence_after();

#pragma unroll 1
  for (int c = 0; c < NSLABS; ++c) {
    const int col0 = n_base + c * 32;
    if (col0 >= p.N) break;  // warp-uniform
    if constexpr (MODE != EPI_GENERIC && MODE != EPI_MAP) {
      if (c + 1 < NSLABS) epi_prefetch<MODE>(p, nxt, row_base, col0 + 32 + j4 * 4, first_split, sub);
    }
    // ---- phase A
    {
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr0 + static_cast<uint32_t>(c * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128(wr_addr + ((static_cast<uint32_t>(j) ^ wr_sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2],
               v[4 * j + 3]);
    }
    __syncwarp();
    // ---- phase B
    const int col = col0 + j4 * 4;
    const bool col_ok = col < p.N;  // N % 8 == 0: a 4-column chunk is all-in or all-out
    float4 acc[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rl = it * 4 + sub;
      acc[it] = lds128(stage_addr + rl * 128 +
                       ((static_cast<uint32_t>(j4) ^ static_cast<uint32_t>(rl & 7)) << 4));
    }
    if constexpr (MODE == EPI_GENERIC) {
#pragma unroll 1
      for (int it = 0; it < 8; ++it) {
        const int row = row_base + it * 4 + sub;
        if (row < p.M && col_ok) epi_generic_chunk(p, acc[it], row, col, first_split);
      }
    } else if constexpr (MODE == EPI_MAP) {
      // per-slab column part of the output index (un-patchify: col -> (channel, py, px))
      long long ocol = col;
      if (unpatch) {
        const int pp = p.up_ph * p.up_pw;
        const int ch = col / pp, rr2 = col - ch * pp;
        const int py = rr2 / p.up_pw, px = rr2 - py * p.up_pw;
        ocol = (static_cast<long long>(ch) * (p.up_gh * p.up_ph) + py) * (static_cast<long long>(p.up_gw) * p.up_pw) + px;
      }
      float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias != nullptr && col_ok) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = row_base + it * 4 + sub;
        if (row < p.M && col_ok) {
          float f0 = acc[it].x + bias4.x, f1 = acc[it].y + bias4.y, f2 = acc[it].z + bias4.z,
                f3 = acc[it].w + bias4.w;
          if (p.residual != nullptr) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(p.residual + map_res[it] * p.ld_res + col));
            f0 += r.x; f1 += r.y; f2 += r.z; f3 += r.w;
          }
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + map_out[it] + ocol) =
              make_float4(f0, f1, f2, f3);
        }
      }
    } else {
      const long long row0 = row_base + sub;
      constexpr int OUT_BYTES = (epi_has_res(MODE) || MODE == EPI_F32) ? 4 : 2;
      uint8_t* optr = reinterpret_cast<uint8_t*>(p.out) + (row0 * p.ldc + col) * OUT_BYTES;
      const long long ostep = 4 * p.ldc * OUT_BYTES;
      __nv_bfloat16* aptr = nullptr;
      long long astep = 0;
      if constexpr (MODE == EPI_GELU) {
        if (p.aux_out) {
          aptr = p.aux_out + row0 * p.ld_aux + col;
          astep = 4 * p.ld_aux;
        }
      }
      const bool atomic = (MODE == EPI_F32) && (p.epilogue & MB_EPI_ATOMIC);
      [[maybe_unused]] float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f, cs3 = 0.f;  // EPI_DGELU: column sums of this thread's rows
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = row_base + it * 4 + sub;
        float f0, f1, f2, f3;
        if constexpr (epi_ln_in(MODE)) {
          f0 = fmaf(acc[it].x, ln_rstd[it], fmaf(-ln_mr[it], cur.c1.x, cur.bias.x));
          f1 = fmaf(acc[it].y, ln_rstd[it], fmaf(-ln_mr[it], cur.c1.y, cur.bias.y));
          f2 = fmaf(acc[it].z, ln_rstd[it], fmaf(-ln_mr[it], cur.c1.z, cur.bias.z));
          f3 = fmaf(acc[it].w, ln_rstd[it], fmaf(-ln_mr[it], cur.c1.w, cur.bias.w));
        } else {
          f0 = acc[it].x + cur.bias.x; f1 = acc[it].y + cur.bias.y;
          f2 = acc[it].z + cur.bias.z; f3 = acc[it].w + cur.bias.w;
        }
        const bool ok = (row < p.M) && col_ok;
        if constexpr (MODE == EPI_GELU_LN) {
          gelu_fast2(f0, f1); gelu_fast2(f2, f3);
        }
        if constexpr (MODE == EPI_GELU) {
          if (aptr != nullptr && ok) {
            uint2 pk;
            pk.x = pack_bf16x2(f0, f1);
            pk.y = pack_bf16x2(f2, f3);
            *reinterpret_cast<uint2*>(aptr + it * astep) = pk;
          }
          gelu_fast2(f0, f1); gelu_fast2(f2, f3);
        }
        if constexpr (MODE == EPI_DGELU) {
          float2 h0 = unpack_bf16x2(cur.aux[it].x), h1 = unpack_bf16x2(cur.aux[it].y);
          gelu_fast_grad2(h0.x, h0.y);
          gelu_fast_grad2(h1.x, h1.y);
          f0 *= h0.x; f1 *= h0.y; f2 *= h1.x; f3 *= h1.y;
          if (ok) { cs0 += f0; cs1 += f1; cs2 += f2; cs3 += f3; }
        }
        if constexpr (epi_has_res(MODE)) {
          f0 += cur.res[it].x; f1 += cur.res[it].y; f2 += cur.res[it].z; f3 += cur.res[it].w;
        }
        if constexpr (MODE == EPI_RES_LN) {
          if (ok) {
            rs1[it] += (f0 + f1) + (f2 + f3);
            rs2[it] += (f0 * f0 + f1 * f1) + (f2 * f2 + f3 * f3);
            uint2 tw;
            tw.x = pack_bf16x2(f0, f1);
            tw.y = pack_bf16x2(f2, f3);
            *reinterpret_cast<uint2*>(p.twin_out + static_cast<long long>(row) * p.ld_twin + col) = tw;
          }
        }
        if (ok) {
          if constexpr (OUT_BYTES == 4) {
            float* o = reinterpret_cast<float*>(optr + it * ostep);
            if (atomic) {
              red_add_v4(o, f0, f1, f2, f3);
            } else {
              *reinterpret_cast<float4*>(o) = make_float4(f0, f1, f2, f3);
            }
          } else {
            uint2 pk;
            pk.x = pack_bf16x2(f0, f1);
            pk.y = pack_bf16x2(f2, f3);
            *reinterpret_cast<uint2*>(optr + it * ostep) = pk;
          }
        }
      }
      if constexpr (MODE == EPI_DGELU) {
        if (p.colsum_out != nullptr) {  // fold the 4 row groups (lanes j4, j4+8, j4+16, j4+24), one vector reduction
          cs0 += __shfl_xor_sync(0xffffffffu, cs0, 8);  cs1 += __shfl_xor_sync(0xffffffffu, cs1, 8);
          cs2 += __shfl_xor_sync(0xffffffffu, cs2, 8);  cs3 += __shfl_xor_sync(0xffffffffu, cs3, 8);
          cs0 += __shfl_xor_sync(0xffffffffu, cs0, 16); cs1 += __shfl_xor_sync(0xffffffffu, cs1, 16);
          cs2 += __shfl_xor_sync(0xffffffffu, cs2, 16); cs3 += __shfl_xor_sync(0xffffffffu, cs3, 16);
          if (sub == 0 && col_ok) red_add_v4(p.colsum_out + col, cs0, cs1, cs2, cs3);
        }
      }
      cur = nxt;
    }
    __syncwarp();  // the slab is rewritten by the next phase A
  }
  if constexpr (MODE == EPI_RES_LN) {
    // the 8 lanes that share a row (j4 = 0..7) fold their partial sums; one reduction pair per row and warp
#pragma unroll
    for (int it = 0; it < 8; ++it) {
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        rs1[it] += __shfl_xor_sync(0xffffffffu, rs1[it], o);
        rs2[it] += __shfl_xor_sync(0xffffffffu, rs2[it], o);
      }
      const int row = row_base + it * 4 + sub;
      const int part = n_base / (NSLABS * 32);   // this warp's column range = one slot of the row's statistics
      if (j4 == 0 && row < p.M && part < p.ln_parts)
        *reinterpret_cast<float2*>(p.row_stats + 2ll * (static_cast<long long>(row) * (p.ln_parts + 1) + 1 + part)) =
            make_float2(rs1[it], rs2[it]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// shared pieces of the two kernels
// ---------------------------------------------------------------------------------------------
struct TileCoord {
  int split, m_blk, n_blk, kb0, kb1;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmDev& p, int tile) {
  TileCoord t;
  const int tiles_mn = p.m_tiles * p.n_tiles;
  t.split = tile / tiles_mn;
  const int mn = tile - t.split * tiles_mn;
  const int mb = mn / p.n_tiles;
  t.n_blk = mn - mb * p.n_tiles;
  t.m_blk = p.m_reverse ? p.m_tiles - 1 - mb : mb;
  t.kb0 = t.split * p.kb_per_split;
  t.kb1 = min(t.kb0 + p.kb_per_split, p.k_blocks);
  return t;
}

// ---------------------------------------------------------------------------------------------
// one CTA per 128 x BN tile
// ---------------------------------------------------------------------------------------------
template <int BN, int A_LAYOUT, int B_MN, int ESIZE, int EPI>
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
  pdl_wait();  // everything above is private to the CTA; global memory is only touched below

  // The producer and the MMA issuer run as WHOLE, converged warps on warp-uniform values (loop counters, kernel
  // parameters, shuffled registers) and elect one lane only around the TMA / MMA / commit instructions themselves.
  // Under the previous `lane == 0` guard ptxas could not prove the operands uniform and wrapped every UTCHMMA and
  // UTMALDG in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop: ~100 clk of serial issue per MMA (clock64 timeline of
  // the attention issuer, profiles/r02_ncu_attn_fwd4_summary.txt), i.e. ~400 clk per 64-wide K block against the
  // 512 clk the four MMAs of that block take on the tensor core -- the issuer was nearly co-critical.
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int m0 = t.m_blk * kBM;
      const int n0 = t.n_blk * BN;
      for (int kb = t.kb0; kb < t.kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        if (leader) {
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
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
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    pdl_trigger();  // last operand tile requested: the successor may start its prologue while this CTA drains
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const bool leader = elect_one();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
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
      const TileCoord t = decode_tile(p, tile);
      mbar_wait(&tmem_empty_bar[acc_stage], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_u + acc_stage * BN;
      for (int kb = t.kb0; kb < t.kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t b_addr = a_addr + Cfg::kABytes;
        // K step k = start address + k * kstep bytes = + k * (kstep >> 4) in the descriptor's address field
        const uint64_t da0 = make_smem_desc(a_addr, a_lbo, 1024);
        const uint64_t db0 = make_smem_desc(b_addr, b_lbo, 1024);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t acc = (kb > t.kb0 || k > 0) ? 1u : 0u;
            if constexpr (ESIZE == 2)
              umma_f16_ss(tmem_d, da0 + k * (a_kstep >> 4), db0 + k * (b_kstep >> 4), idesc, acc);
            else
              umma_tf32_ss(tmem_d, da0 + k * (a_kstep >> 4), db0 + k * (b_kstep >> 4), idesc, acc);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (leader) umma_commit(&tmem_full_bar[acc_stage]);  // accumulator complete -> epilogue
      __syncwarp();
      acc_stage ^= 1;
      if (acc_stage == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;  // which half of the BN columns
    constexpr int HALF_COLS = BN / 2;
    const uint32_t stage_addr =
        smem_u32(smem + STAGES * Cfg::kStageBytes + Cfg::kBarBytes + (warp - 4) * 4096);
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(acc_stage * BN + half * HALF_COLS);
      epilogue_warp<EPI, HALF_COLS / 32>(p, stage_addr, taddr, t.m_blk * kBM + q * 32,
                                         t.n_blk * BN + half * HALF_COLS, t.split, lane,
                                         &tmem_full_bar[acc_stage], acc_phase);
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
// CTA pair (cta_group::2) per 256 x BN tile
//   * both CTAs run a TMA producer; completion bytes of BOTH are credited to the leader's full barrier
//   * only the leader (cluster rank 0) issues MMAs; tcgen05.commit multicasts the arrive to the empty /
//     tmem_full barriers of both CTAs
//   * both CTAs' epilogue warps arrive (remotely for rank 1) on the leader's tmem_empty barrier
// ---------------------------------------------------------------------------------------------
template <int BN, int A_LAYOUT, int B_MN, int ESIZE, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const GemmDev p) {
  using Cfg = PairCfg<BN>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int BK = 128 / ESIZE;
  constexpr bool A_MN = (A_LAYOUT == MB_MAJOR_MN);
  static_assert(!(A_MN || B_MN) || ESIZE == 2, "MN-major operands are bf16 only");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                         // used in the leader CTA
  uint64_t* empty_bar = bars + STAGES;               // both CTAs (multicast commit)
  uint64_t* tmem_full_bar = bars + 2 * STAGES;       // both CTAs (multicast commit)
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;  // leader CTA, 16 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

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
      mbar_init(&tmem_empty_bar[s], 16);  // 8 epilogue warps in each CTA of the pair
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above is private to the CTA; global memory is only touched below

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    // (whole warp, converged, uniform operands -- see gemm_kernel above)
    const bool leader = elect_one();
    const uint32_t cta_u = __shfl_sync(0xffffffffu, cta, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < p.total_tiles; tile += num_clusters) {
      const TileCoord t = decode_tile(p, tile);  // m_tiles counts 256-row tiles here
      const int m0 = t.m_blk * 256 + static_cast<int>(cta_u) * kBM;
      const int n0 = t.n_blk * BN + static_cast<int>(cta_u) * (BN / 2);
      for (int kb = t.kb0; kb < t.kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const uint32_t fb = map_to_cta(smem_u32(&full_bar[stage]), 0);
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        if (leader) {
        if (cta_u == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
        if constexpr (A_LAYOUT == MB_MAJOR_K) {
          tma_load_2d_pair(sa, &tma_a, fb, kb * BK, m0);
        } else if constexpr (A_LAYOUT == MB_MAJOR_MN) {
#pragma unroll
          for (int s = 0; s < kBM / 64; ++s)
            tma_load_2d_pair(sa + s * 8192, &tma_a, fb, m0 + s * 64, kb * 64);
        } else {
          const int img = m0 / p.rows_per_img;
          const int nh0 = (m0 - img * p.rows_per_img) / p.grid_w;
          tma_load_5d_pair(sa, &tma_a, fb, 0, 0, kb, nh0, img);
        }
        if constexpr (B_MN == 0) {
          tma_load_2d_pair(sb, &tma_b, fb, kb * BK, n0);
        } else {
#pragma unroll
          for (int s = 0; s < BN / 128; ++s)
            tma_load_2d_pair(sb + s * 8192, &tma_b, fb, n0 + s * 64, kb * 64);
        }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    pdl_trigger();
  } else if (warp == 1 && cta == 0) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    const bool leader = elect_one();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    constexpr uint32_t idesc =
        make_idesc(256, BN, ESIZE == 2 ? kFmtBF16 : kFmtTF32, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
    constexpr uint32_t a_lbo = A_MN ? 8192u : 0u;
    constexpr uint32_t b_lbo = B_MN ? 8192u : 0u;
    constexpr uint32_t a_kstep = A_MN ? 2048u : 32u;
    constexpr uint32_t b_kstep = B_MN ? 2048u : 32u;
    int stage = 0;
    uint32_t phase = 0;
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < p.total_tiles; tile += num_clusters) {
      const TileCoord t = decode_tile(p, tile);
      mbar_wait(&tmem_empty_bar[acc_stage], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_u + acc_stage * BN;
      for (int kb = t.kb0; kb < t.kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t b_addr = a_addr + Cfg::kABytes;
        const uint64_t da0 = make_smem_desc(a_addr, a_lbo, 1024);
        const uint64_t db0 = make_smem_desc(b_addr, b_lbo, 1024);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t acc = (kb > t.kb0 || k > 0) ? 1u : 0u;
            if constexpr (ESIZE == 2)
              umma_f16_ss_pair(tmem_d, da0 + k * (a_kstep >> 4), db0 + k * (b_kstep >> 4), idesc, acc);
            else
              umma_tf32_ss_pair(tmem_d, da0 + k * (a_kstep >> 4), db0 + k * (b_kstep >> 4), idesc, acc);
          }
          umma_commit_pair(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (leader) umma_commit_pair(&tmem_full_bar[acc_stage]);
      __syncwarp();
      acc_stage ^= 1;
      if (acc_stage == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (both CTAs)
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    constexpr int HALF_COLS = BN / 2;
    const uint32_t stage_addr =
        smem_u32(smem + STAGES * Cfg::kStageBytes + Cfg::kBarBytes + (warp - 4) * 4096);
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < p.total_tiles; tile += num_clusters) {
      const TileCoord t = decode_tile(p, tile);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(acc_stage * BN + half * HALF_COLS);
      epilogue_warp<EPI, HALF_COLS / 32>(p, stage_addr, taddr,
                                         t.m_blk * 256 + static_cast<int>(cta) * kBM + q * 32,
                                         t.n_blk * BN + half * HALF_COLS, t.split, lane,
                                         &tmem_full_bar[acc_stage], acc_phase);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&tmem_empty_bar[acc_stage]), 0));
      acc_stage ^= 1;
      if (acc_stage == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// launchers (instantiated in gemm_single.cu / gemm_pair.cu)
// ---------------------------------------------------------------------------------------------
template <int BN, int A_LAYOUT, int B_MN, int ESIZE, int EPI>
int launch_gemm_single(const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_kernel<BN, A_LAYOUT, B_MN, ESIZE, EPI>;
  static PerDeviceOnce configured;
  if (configured.first()) {
    MB_CHECK_CUDA(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  }
  const int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
  MB_CHECK_CUDA(launch_k(kern, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, stream, ta, tb, p));
  return 0;
}

template <int BN, int A_LAYOUT, int B_MN, int ESIZE, int EPI>
int launch_gemm_pair(const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p,
                     cudaStream_t stream) {
  using Cfg = PairCfg<BN>;
  auto kern = gemm_pair_kernel<BN, A_LAYOUT, B_MN, ESIZE, EPI>;
  static PerDeviceOnce configured;
  if (configured.first()) {
    MB_CHECK_CUDA(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  }
  const int max_clusters = sm_count() / 2;
  const int clusters = p.total_tiles < max_clusters ? p.total_tiles : max_clusters;
  MB_CHECK_CUDA(launch_k(kern, dim3(2 * clusters), dim3(kGemmThreads), Cfg::kSmemBytes, stream, ta, tb, p));
  return 0;
}

// layout classes the dispatchers understand
enum : int { LAY_KK_BF16 = 0, LAY_KMN_BF16 = 1, LAY_MNMN_BF16 = 2, LAY_PATCH_TF32 = 3, LAY_KK_TF32 = 4 };

// returns 1 when (layout, epilogue) has no specialised instantiation: the caller retries with EPI_GENERIC
int dispatch_gemm_single(int bn, int layout, int epi, const CUtensorMap& ta, const CUtensorMap& tb,
                         const GemmDev& p, cudaStream_t stream);
int dispatch_gemm_pair(int bn, int layout, int epi, const CUtensorMap& ta, const CUtensorMap& tb,
                       const GemmDev& p, cudaStream_t stream);
// folded-LayerNorm epilogues (EPI_RES_LN / EPI_BF16_LN / EPI_GELU_LN), K-major bf16 operands only (gemm_ln.cu)
int dispatch_gemm_ln(int bn, bool pair, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p,
                     cudaStream_t stream);

}  // namespace mb200
