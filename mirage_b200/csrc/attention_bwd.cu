// mirage_b200/csrc/attention_bwd.cu
//
// Attention backward on tcgen05/TMEM for sm_100a (no mask, no dropout), given the forward's
// log-sum-exp:   P = exp(S*scale - LSE),  dV = P^T dO,  dP = dO V^T,  dS = P o (dP - D) * scale,
//                dQ = dS K,  dK = dS^T Q,   D_i = rowsum(dO_i o O_i).
// Backward of F.scaled_dot_product_attention at mirage/utils.py:181-185 / :216-220.
//
// One CTA per (batch, head).  Outer loop over key blocks j (128 keys), inner loop over query tiles i
// (128 rows).  dK_j / dV_j accumulate in TMEM across the inner loop; dQ_i is produced per (j, i) and
// (when there is more than one key block) accumulated in an fp32 workspace by the thread that owns
// the row -- no atomics anywhere.
//
//   warp 0     TMA producer: K_j, V_j per key block; Q_i, dO_i per (j, i)
//   warp 1     TMEM allocator + MMA issuer (one lane)
//   warps 2-5  one thread per query row: P and dS (bf16) into 128B-swizzled smem, dQ/dK/dV write-out
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

constexpr int kAttnBwdThreads = 192;

template <int HD>
struct AttnBwdCfg {
  static constexpr int kRowBytes = HD * 2;
  static constexpr int kTileBytes = 128 * kRowBytes;
  static constexpr int kPBytes = 128 * 128 * 2;
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kOffK + kTileBytes;
  static constexpr int kOffQ = kOffV + kTileBytes;
  static constexpr int kOffDO = kOffQ + kTileBytes;
  static constexpr int kOffP = kOffDO + kTileBytes;
  static constexpr int kOffDS = kOffP + kPBytes;
  static constexpr int kOffBar = kOffDS + kPBytes;
  static constexpr int kSmemBytes = kOffBar + 128;
  static constexpr int kSwizzle = (HD == 64) ? 128 : 64;
};

template <int HD>
__global__ void __launch_bounds__(kAttnBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                const AttnBwdDev p) {
  using Cfg = AttnBwdCfg<HD>;
  constexpr uint64_t kSw = (HD == 64) ? kDescSwizzle128B : kDescSwizzle64B;
  constexpr uint32_t kSbo = 8 * Cfg::kRowBytes;
  constexpr uint32_t kRowStep16 = 16 * Cfg::kRowBytes;  // 16 rows of an operand tile

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* kv_full = bars + 0;
  uint64_t* kv_empty = bars + 1;
  uint64_t* qdo_full = bars + 2;
  uint64_t* qdo_empty = bars + 3;
  uint64_t* sdp_full = bars + 4;
  uint64_t* pds_full = bars + 5;
  uint64_t* mma2_done = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.x % p.H;
  const int b = blockIdx.x / p.H;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("mirage_b200: attention-bwd smem base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    mbar_init(qdo_full, 1);
    mbar_init(qdo_empty, 1);
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 128);
    mbar_init(mma2_done, 1);
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
  const int kvb = p.kv_blocks, qt = p.q_tiles;

  if (warp == 0 && lane == 0) {
    // ---------------------------------------------------------------- TMA producer
    int it = 0;
    for (int j = 0; j < kvb; ++j) {
      mbar_wait(kv_empty, (j & 1) ^ 1);
      mbar_arrive_expect_tx(kv_full, 2 * Cfg::kTileBytes);
      tma_load_3d(smem + Cfg::kOffK, &tm_k, kv_full, h * HD, j * 128, b);
      tma_load_3d(smem + Cfg::kOffV, &tm_v, kv_full, h * HD, j * 128, b);
      for (int i = 0; i < qt; ++i, ++it) {
        mbar_wait(qdo_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(qdo_full, 2 * Cfg::kTileBytes);
        tma_load_3d(smem + Cfg::kOffQ, &tm_q, qdo_full, h * HD, i * 128, b);
        tma_load_3d(smem + Cfg::kOffDO, &tm_do, qdo_full, h * HD, i * 128, b);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t k_addr = smem_u32(smem + Cfg::kOffK);
    const uint32_t v_addr = smem_u32(smem + Cfg::kOffV);
    const uint32_t q_addr = smem_u32(smem + Cfg::kOffQ);
    const uint32_t do_addr = smem_u32(smem + Cfg::kOffDO);
    const uint32_t p_addr = smem_u32(smem + Cfg::kOffP);
    const uint32_t ds_addr = smem_u32(smem + Cfg::kOffDS);
    constexpr uint32_t idesc_dkv = make_idesc(128, HD, kFmtBF16, 1, 1);
    constexpr uint32_t idesc_dq = make_idesc(128, HD, kFmtBF16, 0, 1);
    int it = 0;
    for (int j = 0; j < kvb; ++j) {
      const int valid = min(128, p.Nk - j * 128);
      const uint32_t ncols = static_cast<uint32_t>((valid + 15) & ~15);
      const uint32_t idesc_s = make_idesc(128, ncols, kFmtBF16, 0, 0);
      mbar_wait(kv_full, j & 1);
      for (int i = 0; i < qt; ++i, ++it) {
        mbar_wait(qdo_full, it & 1);
        tc_fence_after();
        // S = Q_i K_j^T ; dP = dO_i V_j^T      (K-major operands, N = ncols)
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          umma_f16_ss(tmem_base + 0, make_smem_desc(q_addr + k * 32, 0, kSbo, kSw),
                      make_smem_desc(k_addr + k * 32, 0, kSbo, kSw), idesc_s, k > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          umma_f16_ss(tmem_base + 128, make_smem_desc(do_addr + k * 32, 0, kSbo, kSw),
                      make_smem_desc(v_addr + k * 32, 0, kSbo, kSw), idesc_s, k > 0 ? 1u : 0u);
        }
        umma_commit(sdp_full);

        mbar_wait(pds_full, it & 1);
        tc_fence_after();
        // dV_j += P^T dO_i ; dK_j += dS^T Q_i   (A = P / dS read MN-major: M = keys, K = query rows)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          umma_f16_ss(tmem_base + 256, make_smem_desc(p_addr + kk * 2048, 16384, 1024),
                      make_smem_desc(do_addr + kk * kRowStep16, 0, kSbo, kSw), idesc_dkv,
                      (i > 0 || kk > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          umma_f16_ss(tmem_base + 320, make_smem_desc(ds_addr + kk * 2048, 16384, 1024),
                      make_smem_desc(q_addr + kk * kRowStep16, 0, kSbo, kSw), idesc_dkv,
                      (i > 0 || kk > 0) ? 1u : 0u);
        }
        // dQ_i(j) = dS K_j                      (A = dS K-major: M = query rows, K = keys)
        const int ksteps = (valid + 15) >> 4;
        for (int kk = 0; kk < ksteps; ++kk) {
          umma_f16_ss(tmem_base + 384,
                      make_smem_desc(ds_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024),
                      make_smem_desc(k_addr + kk * kRowStep16, 0, kSbo, kSw), idesc_dq,
                      kk > 0 ? 1u : 0u);
        }
        umma_commit(mma2_done);
        umma_commit(qdo_empty);
        if (i == qt - 1) umma_commit(kv_empty);
      }
    }
  } else if (warp >= 2) {
    // ---------------------------------------------------------------- compute warps
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    uint8_t* p_row = smem + Cfg::kOffP + r * 128;
    uint8_t* ds_row = smem + Cfg::kOffDS + r * 128;
    const float log2e = 1.4426950408889634f;
    int it = 0;
    for (int j = 0; j < kvb; ++j) {
      const int valid = min(128, p.Nk - j * 128);
      for (int i = 0; i < qt; ++i, ++it) {
        const int qrow = i * 128 + r;
        const bool row_ok = qrow < p.Nq;
        // per-row statistics: LSE and D = <dO, O>
        float lse2 = 0.f, dsum = 0.f;
        if (row_ok) {
          lse2 = p.lse[(static_cast<long long>(b) * p.H + h) * p.Nq + qrow] * log2e;
          const __nv_bfloat16* orow = p.o + (static_cast<long long>(b) * p.Nq + qrow) * p.ldo + h * HD;
          const __nv_bfloat16* drow = p.d_o + (static_cast<long long>(b) * p.Nq + qrow) * p.lddo + h * HD;
#pragma unroll
          for (int c = 0; c < HD / 8; ++c) {
            const uint4 a = *reinterpret_cast<const uint4*>(orow + c * 8);
            const uint4 g = *reinterpret_cast<const uint4*>(drow + c * 8);
            const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z),
                         a3 = unpack_bf16x2(a.w);
            const float2 g0 = unpack_bf16x2(g.x), g1 = unpack_bf16x2(g.y), g2 = unpack_bf16x2(g.z),
                         g3 = unpack_bf16x2(g.w);
            dsum += a0.x * g0.x + a0.y * g0.y + a1.x * g1.x + a1.y * g1.y + a2.x * g2.x + a2.y * g2.y +
                    a3.x * g3.x + a3.y * g3.y;
          }
        }
        mbar_wait(sdp_full, it & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t vs[32], vd[32];
          if (c * 32 < valid) {  // warp-uniform
            tmem_ld_32x32b_x32(tmem_base + lane_off + c * 32, vs);
            tmem_ld_32x32b_x32(tmem_base + lane_off + 128 + c * 32, vd);
            tmem_ld_wait();
          }
          uint8_t* pd = p_row + (c >> 1) * 16384;
          uint8_t* dd = ds_row + (c >> 1) * 16384;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float pv[8], dv_[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int col = c * 32 + t * 8 + e;
              float pe = 0.f, de = 0.f;
              if (row_ok && col < valid) {
                pe = fast_exp2(fmaf(__uint_as_float(vs[t * 8 + e]), p.scale_log2, -lse2));
                de = pe * (__uint_as_float(vd[t * 8 + e]) - dsum) * p.scale;
              }
              pv[e] = pe;
              dv_[e] = de;
            }
            uint4 pk, dk4;
            pk.x = pack_bf16x2(pv[0], pv[1]);
            pk.y = pack_bf16x2(pv[2], pv[3]);
            pk.z = pack_bf16x2(pv[4], pv[5]);
            pk.w = pack_bf16x2(pv[6], pv[7]);
            dk4.x = pack_bf16x2(dv_[0], dv_[1]);
            dk4.y = pack_bf16x2(dv_[2], dv_[3]);
            dk4.z = pack_bf16x2(dv_[4], dv_[5]);
            dk4.w = pack_bf16x2(dv_[6], dv_[7]);
            const uint32_t c8 = static_cast<uint32_t>((c & 1) * 4 + t);
            *reinterpret_cast<uint4*>(pd + ((c8 ^ sw) << 4)) = pk;
            *reinterpret_cast<uint4*>(dd + ((c8 ^ sw) << 4)) = dk4;
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(pds_full);

        // dQ_i (+)= dS K_j
        mbar_wait(mma2_done, it & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + lane_off + 384 + c * 32, v);
          tmem_ld_wait();
          if (row_ok) {
            const long long grow = static_cast<long long>(b) * p.Nq + qrow;
            if (kvb > 1) {
              float* acc = p.dq_acc + grow * (static_cast<long long>(p.H) * HD) + h * HD + c * 32;
              if (j > 0) {
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                  const float4 a = *reinterpret_cast<const float4*>(acc + e);
                  v[e + 0] = __float_as_uint(__uint_as_float(v[e + 0]) + a.x);
                  v[e + 1] = __float_as_uint(__uint_as_float(v[e + 1]) + a.y);
                  v[e + 2] = __float_as_uint(__uint_as_float(v[e + 2]) + a.z);
                  v[e + 3] = __float_as_uint(__uint_as_float(v[e + 3]) + a.w);
                }
              }
              if (j < kvb - 1) {
#pragma unroll
                for (int e = 0; e < 32; e += 4)
                  *reinterpret_cast<float4*>(acc + e) =
                      make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                  __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
              }
            }
            if (j == kvb - 1) {
              __nv_bfloat16* dst = p.dq + grow * p.lddq + h * HD + c * 32;
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                uint4 pk;
                pk.x = pack_bf16x2(__uint_as_float(v[t * 8 + 0]), __uint_as_float(v[t * 8 + 1]));
                pk.y = pack_bf16x2(__uint_as_float(v[t * 8 + 2]), __uint_as_float(v[t * 8 + 3]));
                pk.z = pack_bf16x2(__uint_as_float(v[t * 8 + 4]), __uint_as_float(v[t * 8 + 5]));
                pk.w = pack_bf16x2(__uint_as_float(v[t * 8 + 6]), __uint_as_float(v[t * 8 + 7]));
                *reinterpret_cast<uint4*>(dst + t * 8) = pk;
              }
            }
          }
        }
      }
      // dK_j, dV_j complete (the last mma2_done of this key block has been waited on above)
      const int krow = j * 128 + r;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        __nv_bfloat16* base = which == 0 ? p.dv : p.dk;
        const long long ld = which == 0 ? p.lddv : p.lddk;
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + lane_off + 256 + which * 64 + c * 32, v);
          tmem_ld_wait();
          if (krow < p.Nk) {
            __nv_bfloat16* dst = base + (static_cast<long long>(b) * p.Nk + krow) * ld + h * HD + c * 32;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              uint4 pk;
              pk.x = pack_bf16x2(__uint_as_float(v[t * 8 + 0]), __uint_as_float(v[t * 8 + 1]));
              pk.y = pack_bf16x2(__uint_as_float(v[t * 8 + 2]), __uint_as_float(v[t * 8 + 3]));
              pk.z = pack_bf16x2(__uint_as_float(v[t * 8 + 4]), __uint_as_float(v[t * 8 + 5]));
              pk.w = pack_bf16x2(__uint_as_float(v[t * 8 + 6]), __uint_as_float(v[t * 8 + 7]));
              *reinterpret_cast<uint4*>(dst + t * 8) = pk;
            }
          }
        }
      }
      tc_fence_before();
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
  CUtensorMap tq, tk, tv, tdo;
  uint32_t box[3] = {(uint32_t)HD, 128u, 1u};
  {
    uint64_t dims[3] = {(uint64_t)a->heads * HD, (uint64_t)a->nq, (uint64_t)a->batch};
    uint64_t s1[2] = {(uint64_t)a->ldq * 2, (uint64_t)a->nq * a->ldq * 2};
    if (make_tensor_map(&tq, a->q, kTmaBF16, 3, dims, s1, box, Cfg::kSwizzle)) return -1;
    uint64_t s2[2] = {(uint64_t)a->lddo * 2, (uint64_t)a->nq * a->lddo * 2};
    if (make_tensor_map(&tdo, a->d_out, kTmaBF16, 3, dims, s2, box, Cfg::kSwizzle)) return -1;
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
  static bool configured = false;
  if (!configured) {
    MB_CHECK_CUDA(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const long long grid = (long long)p.B * p.H;
  kern<<<(unsigned)grid, kAttnBwdThreads, Cfg::kSmemBytes, stream>>>(tq, tk, tv, tdo, p);
  MB_CHECK_CUDA(cudaGetLastError());
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
