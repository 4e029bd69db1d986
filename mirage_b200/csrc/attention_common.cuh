// mirage_b200/csrc/attention_common.cuh
//
// Pieces shared by the two attention-forward kernels (attention.cu: two query tiles x two column halves;
// attention4.cu: four query tiles, one thread per row): launch parameters, the TS-MMA / bulk-copy wrappers and
// the register-resident exponentiation of one 64-column score row.
#pragma once
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
  //   k_tail keys   -> rank-1 update on CUDA cores (scores in the softmax threads, P_tail V_tail in
  //                    the epilogue warps); the rows travel to smem with the item's Q tiles
  //   q_tail rows   -> attn_tail_rows_kernel (CUDA cores, one CTA per (batch, head)).  Computing that row
//                    inside this kernel (epilogue warps reading the K / V stages from smem) was tried
//                    and was slower: the extra release count on every K / V stage couples the ring to
//                    the slowest warps (0.79-0.85 ms vs 0.69 ms at cfg 2)
  int Nq_main, Nk_main, k_tail;
  int q_tiles, q_pairs, kv_blocks;
  int stagger;  // cycles softmax group 1 idles once at kernel start (anti-phases the two groups)
  int reverse;  // (batch, head) problems are walked from the end (take_direction(), common.cuh); attention4 / _small
  float scale_log2;
};

constexpr int kMaxTail = 4;  // largest remainder (mod 128) that is peeled off instead of padded

constexpr int kAttnThreads = 768;  // 6 warpgroups: {TMA, MMA, MMA, -}, 4 x softmax (group, column half), epilogue
constexpr int kKvStages = 3;
constexpr int kDefaultPoly = 0;
constexpr int kDefaultStagger = 0;
constexpr float kLazyMaxLog2 = 8.0f;  // rescale O only when a row max grew by more than 2^8

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (P, bf16 pairs, one query row per TMEM lane) is read
// from tensor memory; issued by ONE thread.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// plain (non-tensor) bulk copy global -> smem, completion bytes credited to an mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes,
                                          uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// 2^x for a pair of elements on the FMA pipe instead of the MUFU (the softmax is bound by the 16
// ex2/clk/SM of the XU pipe: ncu r01 v4, mio/XU throttle on every MUFU.EX2).  Cody-Waite: n = rint(x)
// through the 1.5*2^23 magic constant, f = x - n in [-0.5, 0.5], 2^f by a cubic (max rel. error 2.1e-4,
// an order of magnitude below the bf16 rounding of P), exponent patched in with one integer IMAD.
__device__ __forceinline__ void exp2_poly2(float& e0, float& e1, float x0, float x1) {
  x0 = fmaxf(x0, -126.f);
  x1 = fmaxf(x1, -126.f);
  constexpr float kMagic = 12582912.f;  // 1.5 * 2^23
  float t0, t1, n0, n1, f0, f1, p0, p1;
  fadd2(t0, t1, x0, x1, kMagic, kMagic);
  fadd2(n0, n1, t0, t1, -kMagic, -kMagic);
  ffma2(f0, f1, n0, n1, -1.f, -1.f, x0, x1);
  ffma2(p0, p1, f0, f1, 0.054850947f, 0.054850947f, 0.24181366f, 0.24181366f);
  ffma2(p0, p1, p0, p1, f0, f1, 0.69324836f, 0.69324836f);
  ffma2(p0, p1, p0, p1, f0, f1, 0.99998821f, 0.99998821f);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// order-pinned (asm volatile) forms for exp_row's three passes
__device__ __forceinline__ void ffma2_v(uint32_t& d0, uint32_t& d1, float b, float c) {
  asm volatile(
      "{\n"
      ".reg .b64 ra, rb, rc;\n"
      "mov.b64 ra, {%0, %1};\n"
      "mov.b64 rb, {%2, %2};\n"
      "mov.b64 rc, {%3, %3};\n"
      "fma.rn.f32x2 ra, ra, rb, rc;\n"
      "mov.b64 {%0, %1}, ra;\n"
      "}\n"
      : "+r"(d0), "+r"(d1)
      : "f"(b), "f"(c));
}

__device__ __forceinline__ void ex2_v(uint32_t& x) {
  asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x));
}

__device__ __forceinline__ void fadd2_v(float& s0, float& s1, uint32_t a0, uint32_t a1) {
  asm volatile(
      "{\n"
      ".reg .b64 ra, rs;\n"
      "mov.b64 ra, {%2, %3};\n"
      "mov.b64 rs, {%0, %1};\n"
      "add.rn.f32x2 rs, rs, ra;\n"
      "mov.b64 {%0, %1}, rs;\n"
      "}\n"
      : "+f"(s0), "+f"(s1)
      : "r"(a0), "r"(a1));
}

// exponentiate one half S row (64 columns) held in registers: sreg[i] (fp32 bits, i < 32*nchunks) ->
// packed bf16 pairs in sreg[i/2]; returns the row sum.  FULL: all 64 columns valid, no masking code.
// Three passes in pinned program order, each a run of mutually independent instructions (the
// compiler's own interleaving left every instruction waiting on its predecessor: ncu r01 v4, 7 clk per
// instruction with 'wait' the top stall):  x = s*scale - m (64 FFMA2);  e = 2^x in place (128 MUFU,
// back to back -- the XU pipe is the only thing that should stall here, and the other softmax warp of
// the SM sub-partition issues into the gaps);  row sum on 4 independent FADD2 chains + bf16 packing.
// POLY: of every 16 element pairs, this many go through exp2_poly2 instead of MUFU.EX2.
template <bool FULL, int POLY>
__device__ __forceinline__ float exp_row(uint32_t (&sreg)[64], float scale_log2, float neg_m, int valid,
                                         int nchunks) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if (FULL || c < nchunks) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) ffma2_v(sreg[c * 32 + i], sreg[c * 32 + i + 1], scale_log2, neg_m);
    }
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if (FULL || c < nchunks) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const int idx = c * 32 + i;
        if (((i >> 1) * POLY) / 16 != (((i >> 1) + 1) * POLY) / 16) {
          float e0, e1;
          exp2_poly2(e0, e1, __uint_as_float(sreg[idx]), __uint_as_float(sreg[idx + 1]));
          sreg[idx] = __float_as_uint(e0);
          sreg[idx + 1] = __float_as_uint(e1);
        } else {
          ex2_v(sreg[idx]);
          ex2_v(sreg[idx + 1]);
        }
        if (!FULL) {
          if (idx >= valid) sreg[idx] = 0u;
          if (idx + 1 >= valid) sreg[idx + 1] = 0u;
        }
      }
    }
  }
  float sum[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) sum[a] = 0.f;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if (FULL || c < nchunks) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const int idx = c * 32 + i;
        const int a = (i >> 1) & 3;
        fadd2_v(sum[2 * a], sum[2 * a + 1], sreg[idx], sreg[idx + 1]);
        sreg[idx >> 1] = pack_bf16x2(__uint_as_float(sreg[idx]), __uint_as_float(sreg[idx + 1]));
      }
    }
  }
  return ((sum[0] + sum[1]) + (sum[2] + sum[3])) + ((sum[4] + sum[5]) + (sum[6] + sum[7]));
}


// launcher of the four-tile kernel (attention4.cu); returns 1 when the problem does not fit its assumptions
int launch_attn_fwd4(const mb_attn_args* a, const AttnDev& p, int poly, cudaStream_t stream);
// launcher of the single-tile kernel (attention_small.cu: head_dim 64, Nq <= 128, Nk <= 128); same convention
int launch_attn_fwd_small(const mb_attn_args* a, const AttnDev& p, cudaStream_t stream);

}  // namespace mb200
