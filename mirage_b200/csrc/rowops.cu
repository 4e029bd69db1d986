// mirage_b200/csrc/rowops.cu
//
// HBM-bound row kernels of the MIRAGE hot path (one warp per token row, 16-byte accesses):
//   LayerNorm forward / backward         nn.LayerNorm(eps=1e-6), mirage/utils.py:260-261
//   visible-token gather / scatter       mirage/model.py:384-391
//   bias gradient (column sums)          backward of every nn.Linear bias
//   fp32 -> bf16 cast, global-token fill
#include <cstdlib>

#include "../../include/mirage_b200.h"
#include "common.cuh"

namespace mb200 {

constexpr int kRowThreads = 256;  // 8 warps = 8 rows per block

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ------------------------------------------------------------------------------------------
// LayerNorm forward: y = (x - mean) * rstd * w + b, statistics in fp32 (two-pass, centred variance).
// VPL = float4 vectors per lane; D = 128 * VPL covers 128..1024 (D % 128 == 0); a generic loop
// version handles any D % 4 == 0.
// ------------------------------------------------------------------------------------------
template <int VPL, bool OUT_BF16>
__global__ void __launch_bounds__(kRowThreads)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ b, void* __restrict__ y, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, long long M, int D, long long ldx, long long ldy,
                     float eps, int reverse) {
  pdl_sync();
  long long row = (long long)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
  if (row >= M) return;
  // `reverse`: CTAs walk the rows from the END.  x was just written by a GEMM that walks its row tiles upwards,
  // so the last ~100 MB of rows are still in the 126 MB L2 when this kernel starts -- and the GEMM that consumes
  // y starts at row 0, i.e. with what this kernel wrote last.
  if (reverse) row = M - 1 - row;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * ldx;
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = ld4(xr + (i * 32 + lane) * 4);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
    q += (a * a + c * c) + (d * d + e * e);
  }
  const float rstd = rsqrtf(warp_sum(q) / D + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int col = (i * 32 + lane) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(w + col));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b + col));
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + bb.x;
    o.y = (v[i].y - mean) * rstd * g.y + bb.y;
    o.z = (v[i].z - mean) * rstd * g.z + bb.z;
    o.w = (v[i].w - mean) * rstd * g.w + bb.w;
    if (OUT_BF16) {
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + row * ldy + col) = pk;
    } else {
      st4(reinterpret_cast<float*>(y) + row * ldy + col, o);
    }
  }
}

// LayerNorm backward.  dy is bf16 (gradient of the GEMM operand) or fp32.
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w
//   optionally dx += dres (the gradient flowing through the residual connection, fp32)
//   optionally dx is ALSO written as bf16 (dx_bf16): the operand of the wgrad / dgrad GEMMs that
//   consume it next, so no separate cast kernel ever re-reads dx.
// Persistent blocks (2 per SM); every warp walks rows with a grid stride and accumulates its
// dw = sum dy * xhat, db = sum dy partials in its own shared-memory slab (registers hold only the row
// being processed, so two blocks fit per SM without spills).
// Per-block partials go to part[blk, 2, D]; a second kernel reduces them (deterministic, no atomics).
// DXSUM: a third per-warp slab accumulates the column sums of dx -- the bias gradient of the Linear layer whose
// output gradient dx is (Attention.proj for the block's second LayerNorm, the previous block's Mlp.fc2 for the
// first) -- so that no separate column-sum pass has to re-read dx.
template <int VPL, bool DY_BF16, bool DXSUM>
__global__ void __launch_bounds__(kRowThreads, 2)
layernorm_bwd_kernel(const void* __restrict__ dy, const float* __restrict__ x,
                     const float* __restrict__ w, const float* __restrict__ mean_in,
                     const float* __restrict__ rstd_in, const float* __restrict__ dres,
                     float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_bf16,
                     float* __restrict__ part, long long M, int D, long long ldx, long long lddy,
                     long long lddx, int reverse) {
  extern __shared__ float sred[];  // [8 warps][NS][D]
  constexpr int NS = DXSUM ? 3 : 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sw = sred + warp * NS * D;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int col = (i * 32 + lane) * 4;
    st4(sw + col, make_float4(0.f, 0.f, 0.f, 0.f));
    st4(sw + D + col, make_float4(0.f, 0.f, 0.f, 0.f));
    if (DXSUM) st4(sw + 2 * D + col, make_float4(0.f, 0.f, 0.f, 0.f));
  }
  const float invD = 1.f / D;
  pdl_sync();
  for (long long r_ = (long long)blockIdx.x * (kRowThreads / 32) + warp; r_ < M;
       r_ += (long long)gridDim.x * (kRowThreads / 32)) {
    const long long row = reverse ? M - 1 - r_ : r_;   // see layernorm_fwd_kernel
    const float mean = mean_in[row], rstd = rstd_in[row];
    float4 xh[VPL], d[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int col = (i * 32 + lane) * 4;
      xh[i] = ld4(x + row * ldx + col);
      if (DY_BF16) {
        const uint2 pk = *reinterpret_cast<const uint2*>(
            reinterpret_cast<const __nv_bfloat16*>(dy) + row * lddy + col);
        const float2 a = unpack_bf16x2(pk.x), c = unpack_bf16x2(pk.y);
        d[i] = make_float4(a.x, a.y, c.x, c.y);
      } else {
        d[i] = ld4(reinterpret_cast<const float*>(dy) + row * lddy + col);
      }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int col = (i * 32 + lane) * 4;
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + col));
      xh[i] = make_float4((xh[i].x - mean) * rstd, (xh[i].y - mean) * rstd, (xh[i].z - mean) * rstd,
                          (xh[i].w - mean) * rstd);
      // parameter-gradient partials (own slab: no conflicts, no atomics)
      float4 a = ld4(sw + col), b = ld4(sw + D + col);
      a.x += d[i].x * xh[i].x; a.y += d[i].y * xh[i].y; a.z += d[i].z * xh[i].z; a.w += d[i].w * xh[i].w;
      b.x += d[i].x; b.y += d[i].y; b.z += d[i].z; b.w += d[i].w;
      st4(sw + col, a);
      st4(sw + D + col, b);
      d[i] = make_float4(d[i].x * wv.x, d[i].y * wv.y, d[i].z * wv.z, d[i].w * wv.w);  // g = dy * w
      s1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      s2 += (d[i].x * xh[i].x + d[i].y * xh[i].y) + (d[i].z * xh[i].z + d[i].w * xh[i].w);
    }
    const float m1 = warp_sum(s1) * invD, m2 = warp_sum(s2) * invD;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int col = (i * 32 + lane) * 4;
      float4 o;
      o.x = rstd * (d[i].x - m1 - xh[i].x * m2);
      o.y = rstd * (d[i].y - m1 - xh[i].y * m2);
      o.z = rstd * (d[i].z - m1 - xh[i].z * m2);
      o.w = rstd * (d[i].w - m1 - xh[i].w * m2);
      if (dres) {
        const float4 r = ld4(dres + row * lddx + col);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      st4(dx + row * lddx + col, o);
      if (DXSUM) {
        float4 c = ld4(sw + 2 * D + col);
        c.x += o.x; c.y += o.y; c.z += o.z; c.w += o.w;
        st4(sw + 2 * D + col, c);
      }
      if (dx_bf16) {
        uint2 pk;
        pk.x = pack_bf16x2(o.x, o.y);
        pk.y = pack_bf16x2(o.z, o.w);
        *reinterpret_cast<uint2*>(dx_bf16 + row * (long long)D + col) = pk;
      }
    }
  }
  // block reduction of the parameter-gradient partials
  __syncthreads();
  for (int c = threadIdx.x; c < NS * D; c += kRowThreads) {
    float acc = 0.f;
#pragma unroll
    for (int wv = 0; wv < kRowThreads / 32; ++wv) acc += sred[wv * NS * D + c];
    part[(long long)blockIdx.x * NS * D + c] = acc;
  }
}

// out[c] (+)= sum_r part[r, c]   (c < cols).  Block = 32 columns x 32 row lanes: the partial rows are
// summed in parallel (a single thread per column walking hundreds of rows is a serial chain of
// dependent-latency loads), then folded through shared memory.  out1 != NULL splits the columns into
// two outputs of D each (LayerNorm dweight | dbias).
__global__ void __launch_bounds__(1024)
colsum_final_kernel(const float* __restrict__ part, float* __restrict__ out0, float* __restrict__ out1,
                    int rows, int cols, int D, int accumulate, float* __restrict__ out2 = nullptr,
                    int accumulate2 = 0) {
  __shared__ float sm[32][33];
  pdl_sync();
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < cols)
    for (int r = threadIdx.y; r < rows; r += 32) acc += part[(long long)r * cols + c];
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += sm[k][threadIdx.x];
    if (out2 != nullptr && c >= 2 * D) {
      float* o = out2 + (c - 2 * D);
      *o = accumulate2 ? (*o + t) : t;
    } else {
      float* o = (out1 == nullptr || c < D) ? (out0 + c) : (out1 + (c - D));
      *o = accumulate ? (*o + t) : t;
    }
  }
}

// ------------------------------------------------------------------------------------------
// column sums of a bf16 / fp32 matrix (bias gradients): part[blk, N] then colsum_final_kernel.
// Thread = one 16-byte column chunk (8 bf16 / 4 fp32) x one row lane; a block is CPB chunks wide and
// 256 / CPB row lanes tall and walks rows_per_block rows with 8 independent 16-byte loads in flight
// per thread; the row lanes are folded through shared memory, one partial row per block.
// ------------------------------------------------------------------------------------------
template <bool IN_BF16>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const void* __restrict__ a, float* __restrict__ part, float* __restrict__ direct,
                      long long M, int N, long long lda, int rows_per_block, int cpb, int reverse) {
  constexpr int EPC = IN_BF16 ? 8 : 4;  // elements per 16-byte chunk
  __shared__ float sm[256 * 8];
  pdl_sync();
  const int lanes = 256 / cpb;
  const int cx = threadIdx.x % cpb, ly = threadIdx.x / cpb;
  const int chunk = blockIdx.y * cpb + cx;
  const int col = chunk * EPC;
  // `reverse`: the first CTAs take the LAST row blocks -- the matrix (the qkv gradient attention-backward has just
  // written upwards: 156 MB at cfg 4, more than the L2 holds) is then read starting with what is still on chip.
  // Row block rb keeps its partial slot, so the result does not depend on the order.
  const int rb = reverse ? static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x);
  const long long row0 = (long long)rb * rows_per_block;
  const long long row1 = (row0 + rows_per_block < M) ? row0 + rows_per_block : M;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (col < N) {
    const uint8_t* base = reinterpret_cast<const uint8_t*>(a) + (size_t)col * (IN_BF16 ? 2 : 4);
    const size_t ldb = (size_t)lda * (IN_BF16 ? 2 : 4);
    long long r = row0 + ly;
    for (; r + 7 * lanes < row1; r += 8 * lanes) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const uint4*>(base + (size_t)(r + u * lanes) * ldb);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (IN_BF16) {
          const float2 f0 = unpack_bf16x2(v[u].x), f1 = unpack_bf16x2(v[u].y), f2 = unpack_bf16x2(v[u].z),
                       f3 = unpack_bf16x2(v[u].w);
          acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y;
          acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
        } else {
          acc[0] += __uint_as_float(v[u].x); acc[1] += __uint_as_float(v[u].y);
          acc[2] += __uint_as_float(v[u].z); acc[3] += __uint_as_float(v[u].w);
        }
      }
    }
    for (; r < row1; r += lanes) {
      const uint4 v = *reinterpret_cast<const uint4*>(base + (size_t)r * ldb);
      if (IN_BF16) {
        const float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y), f2 = unpack_bf16x2(v.z),
                     f3 = unpack_bf16x2(v.w);
        acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y;
        acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
      } else {
        acc[0] += __uint_as_float(v.x); acc[1] += __uint_as_float(v.y);
        acc[2] += __uint_as_float(v.z); acc[3] += __uint_as_float(v.w);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < EPC; ++e) sm[(ly * EPC + e) * cpb + cx] = acc[e];
  __syncthreads();
  if (ly == 0 && col < N) {
    float t[EPC];
#pragma unroll
    for (int e = 0; e < EPC; ++e) {
      t[e] = 0.f;
      for (int l = 0; l < lanes; ++l) t[e] += sm[(l * EPC + e) * cpb + cx];
    }
    if (direct != nullptr) {
      // accumulate mode (gradient buckets): one vector reduction per 4 columns straight into the
      // destination, no partial buffer and no second kernel
#pragma unroll
      for (int e = 0; e < EPC; e += 4) red_add_v4(direct + col + e, t[e], t[e + 1], t[e + 2], t[e + 3]);
    } else {
      float* o = part + (long long)rb * N + col;
#pragma unroll
      for (int e = 0; e < EPC; ++e) o[e] = t[e];
    }
  }
}

// ------------------------------------------------------------------------------------------
// visible-token gather (+ global tokens appended last) and its scatter backward
// ------------------------------------------------------------------------------------------
// out[b, j, :] = src[b, ids[b, j], :]  (j < n_keep);  out[b, n_keep + t, :] = glob[t, :]
__global__ void __launch_bounds__(kRowThreads)
token_gather_kernel(const float* __restrict__ src, const long long* __restrict__ ids,
                    const float* __restrict__ glob, float* __restrict__ out, int B, int n_src,
                    int n_keep, int n_glob, int D) {
  const long long row = (long long)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
  const int n_out = n_keep + n_glob;
  if (row >= (long long)B * n_out) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(row / n_out), j = (int)(row % n_out);
  const float* s;
  if (j < n_keep) {
    long long id = ids[(long long)b * n_keep + j];
    id = id < 0 ? 0 : (id >= n_src ? n_src - 1 : id);  // torch.gather would raise; clamp, never fault
    s = src + ((long long)b * n_src + id) * D;
  } else {
    s = glob + (long long)(j - n_keep) * D;
  }
  float* o = out + row * D;
  for (int c = lane * 4; c < D; c += 128) st4(o + c, ld4(s + c));
}

// dsrc must be zero-filled by the caller.  dsrc[b, ids[b, j], :] = dout[b, j, :]
// (ids are unique within a sample, so a plain store is exact and deterministic.)
__global__ void __launch_bounds__(kRowThreads)
token_scatter_kernel(const float* __restrict__ dout, const long long* __restrict__ ids,
                     float* __restrict__ dsrc, int B, int n_src, int n_keep, int n_glob, int D) {
  const long long row = (long long)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
  if (row >= (long long)B * n_keep) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(row / n_keep), j = (int)(row % n_keep);
  long long id = ids[row];
  if (id < 0 || id >= n_src) return;
  const float* s = dout + ((long long)b * (n_keep + n_glob) + j) * D;
  float* o = dsrc + ((long long)b * n_src + id) * D;
  for (int c = lane * 4; c < D; c += 128) st4(o + c, ld4(s + c));
}

// dglob[t, :] = sum_b dout[b, n_keep + t, :]    (one block per (t, 1024-column slab))
__global__ void global_token_grad_kernel(const float* __restrict__ dout, float* __restrict__ dglob,
                                         int B, int n_keep, int n_glob, int D) {
  const int t = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) acc += dout[((long long)b * (n_keep + n_glob) + n_keep + t) * D + c];
  dglob[(long long)t * D + c] = acc;
}

// out[b, row_off + t, :] = glob[t, :]  -- fills the global-token rows of the un-masked token buffer
__global__ void fill_rows_kernel(const float* __restrict__ glob, float* __restrict__ out, int B,
                                 int n_rows_total, int row_off, int n_glob, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)n_glob * (D / 4);
  if (i >= (long long)B * per) return;
  const int b = (int)(i / per);
  const long long r = i % per;
  const int t = (int)(r / (D / 4)), c = (int)(r % (D / 4)) * 4;
  st4(out + ((long long)b * n_rows_total + row_off + t) * D + c, ld4(glob + (long long)t * D + c));
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                     long long n4) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = ld4(in + i * 4);
  uint2 pk;
  pk.x = pack_bf16x2(v.x, v.y);
  pk.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(out + i * 4) = pk;
}

// row direction of the LayerNorm kernels: downwards (their input was written by a GEMM walking upwards), or, with
// MB_SERPENTINE=1, opposite to whatever the previous kernel did (take_direction(), common.cuh)
static int rows_reversed() { return serpentine_enabled() ? (take_direction() < 0 ? 1 : 0) : 1; }

template <bool OUT_BF16>
static int launch_ln_fwd(const float* x, const float* w, const float* b, void* y, float* mean,
                         float* rstd, long long M, int D, long long ldx, long long ldy, float eps,
                         cudaStream_t st) {
  const unsigned grid = (unsigned)((M + 7) / 8);
#define MB_LN(V)                                                                                   \
  case V:                                                                                          \
    MB_CHECK_CUDA(launch_row(layernorm_fwd_kernel<V, OUT_BF16>, dim3(grid), dim3(kRowThreads), 0, st, x, w, b, \
                           y, mean, rstd, M, D, ldx, ldy, eps, rows_reversed()));                  \
    break;
  switch (D / 128) {
    MB_LN(1) MB_LN(2) MB_LN(3) MB_LN(4) MB_LN(5) MB_LN(6) MB_LN(7) MB_LN(8)
    default:
      MB_REQUIRE(false, "layernorm: D=%d unsupported (multiple of 128, <= 1024)", D);
  }
#undef MB_LN
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mb200

using namespace mb200;

extern "C" {

int mb_layernorm_fwd(const float* x, const float* weight, const float* bias, void* y,
                     int32_t y_dtype, float* mean, float* rstd, int64_t rows, int64_t dim,
                     int64_t ldx, int64_t ldy, float eps, void* stream) {
  MB_REQUIRE(x && weight && bias && y, "mb_layernorm_fwd: null pointer");
  MB_REQUIRE(rows > 0, "mb_layernorm_fwd: rows must be positive");
  MB_REQUIRE(dim % 128 == 0 && dim >= 128 && dim <= 1024,
             "mb_layernorm_fwd: dim=%lld unsupported (multiple of 128, <= 1024)", (long long)dim);
  MB_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0, "mb_layernorm_fwd: leading dims must be multiples of 4");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (y_dtype == MB_BF16)
    return launch_ln_fwd<true>(x, weight, bias, y, mean, rstd, rows, (int)dim, ldx, ldy, eps, st);
  return launch_ln_fwd<false>(x, weight, bias, y, mean, rstd, rows, (int)dim, ldx, ldy, eps, st);
}

static int ln_bwd_blocks(int64_t rows) {
  const int64_t need = (rows + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 2;
  return (int)(need < cap ? need : cap);
}

int64_t mb_layernorm_bwd_workspace(int64_t rows, int64_t dim) {
  return (int64_t)ln_bwd_blocks(rows) * 3 * dim * (int64_t)sizeof(float);
}

int mb_layernorm_bwd(const void* dy, int32_t dy_dtype, const float* x, const float* weight,
                     const float* mean, const float* rstd, const float* dres, float* dx,
                     void* dx_bf16, float* dweight, float* dbias, int32_t accumulate,
                     void* workspace, int64_t rows, int64_t dim, int64_t ldx, int64_t lddy,
                     int64_t lddx, float* dx_colsum, int32_t dx_colsum_accumulate, void* stream) {
  MB_REQUIRE(dy && x && weight && mean && rstd && dx && dweight && dbias && workspace,
             "mb_layernorm_bwd: null pointer");
  MB_REQUIRE(dim % 128 == 0 && dim >= 128 && dim <= 1024,
             "mb_layernorm_bwd: dim=%lld unsupported (multiple of 128, <= 1024)", (long long)dim);
  MB_REQUIRE(rows > 0, "mb_layernorm_bwd: rows must be positive");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)ln_bwd_blocks(rows);
  const int NS = dx_colsum != nullptr ? 3 : 2;
  const size_t smem = (size_t)8 * NS * dim * sizeof(float);
  float* part = reinterpret_cast<float*>(workspace);
  __nv_bfloat16* dxb = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
  const int D = (int)dim;
#define MB_LNB(V, BF)                                                                              \
  if (NS == 3) MB_LNB_(V, BF, true) else MB_LNB_(V, BF, false)
#define MB_LNB_(V, BF, DS)                                                                         \
  {                                                                                                \
    auto kern = layernorm_bwd_kernel<V, BF, DS>;                                                   \
    if (smem > 48 * 1024)                                                                          \
      MB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                         (int)smem));                                              \
    MB_CHECK_CUDA(launch_row(kern, dim3(grid), dim3(kRowThreads), smem, st, dy, x, weight, mean,     \
                           rstd, dres, dx, dxb, part, rows, D, ldx, lddy, lddx, rows_reversed())); \
  }
#define MB_LNB2(V)                                                                                 \
  case V:                                                                                          \
    if (dy_dtype == MB_BF16) MB_LNB(V, true) else MB_LNB(V, false) break;
  switch (D / 128) {
    MB_LNB2(1) MB_LNB2(2) MB_LNB2(3) MB_LNB2(4) MB_LNB2(5) MB_LNB2(6) MB_LNB2(7) MB_LNB2(8)
    default:
      MB_REQUIRE(false, "layernorm_bwd: D=%d unsupported", D);
  }
#undef MB_LNB2
#undef MB_LNB
#undef MB_LNB_
  MB_CHECK_CUDA(cudaGetLastError());
  MB_CHECK_CUDA(launch_row(colsum_final_kernel, dim3((NS * D + 31) / 32), dim3(32, 32), 0, st, part, dweight, dbias,
                         (int)grid, NS * D, D, accumulate, dx_colsum, dx_colsum_accumulate));
  return 0;
}

// tiling of mb_colsum: chunks per block, row blocks
static void colsum_plan(int64_t rows, int64_t cols, int esize, int* cpb, int* grid_y, int* rpb,
                        int* grid_x) {
  const int epc = 16 / esize;
  const int chunks = (int)((cols + epc - 1) / epc);
  int c = 256;
  while (c > 1 && c / 2 >= chunks) c /= 2;  // largest power of two <= 256 that still covers... (>= chunks)
  *cpb = c;
  *grid_y = (chunks + c - 1) / c;
  const int lanes = 256 / c;
  int target = (sm_count() * 4 + *grid_y - 1) / *grid_y;  // ~4 blocks per SM in total
  int64_t r = (rows + target - 1) / target;
  const int64_t min_r = (int64_t)lanes * 8;  // at least one full unrolled batch per thread
  if (r < min_r) r = min_r;
  *rpb = (int)r;
  *grid_x = (int)((rows + r - 1) / r);
}

int64_t mb_colsum_workspace(int64_t rows, int64_t cols) {
  int cpb, gy, rpb, gx_bf16, gx_f32;
  colsum_plan(rows, cols, 2, &cpb, &gy, &rpb, &gx_bf16);
  colsum_plan(rows, cols, 4, &cpb, &gy, &rpb, &gx_f32);
  const int gx = gx_bf16 > gx_f32 ? gx_bf16 : gx_f32;
  return (int64_t)gx * cols * (int64_t)sizeof(float);
}

int mb_colsum(const void* a, int32_t a_dtype, float* out, int32_t accumulate, void* workspace,
              int64_t rows, int64_t cols, int64_t lda, void* stream) {
  MB_REQUIRE(a && out && workspace, "mb_colsum: null pointer");
  MB_REQUIRE(rows > 0 && cols > 0, "mb_colsum: empty matrix");
  const int esize = a_dtype == MB_BF16 ? 2 : 4;
  MB_REQUIRE(cols % (16 / esize) == 0 && lda % (16 / esize) == 0,
             "mb_colsum: cols and lda must be multiples of %d elements (16 bytes)", 16 / esize);
  MB_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0, "mb_colsum: matrix must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int cpb, gy, rpb, gx;
  colsum_plan(rows, cols, esize, &cpb, &gy, &rpb, &gx);
  dim3 grid((unsigned)gx, (unsigned)gy);
  float* part = reinterpret_cast<float*>(workspace);
  // accumulate into a 16-byte aligned destination: single kernel with vector reductions
  float* direct = (accumulate && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? out : nullptr;
  static int rev = -1;   // MB_COLSUM_REVERSE=0: A/B switch
  if (rev < 0) {
    const char* e = getenv("MB_COLSUM_REVERSE");
    rev = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  if (a_dtype == MB_BF16)
    MB_CHECK_CUDA(launch_row(colsum_partial_kernel<true>, grid, dim3(256), 0, st, a, part, direct, rows, (int)cols, lda,
                           rpb, cpb, rev));
  else
    MB_CHECK_CUDA(launch_row(colsum_partial_kernel<false>, grid, dim3(256), 0, st, a, part, direct, rows, (int)cols, lda,
                           rpb, cpb, rev));
  if (direct != nullptr) return 0;
  MB_CHECK_CUDA(launch_row(colsum_final_kernel, dim3((unsigned)((cols + 31) / 32)), dim3(32, 32), 0, st, part, out,
                         nullptr, gx, (int)cols, (int)cols, accumulate, nullptr, 0));
  return 0;
}

int mb_token_gather_fwd(const float* src, const int64_t* ids_keep, const float* global_tokens,
                        float* out, int64_t batch, int64_t n_src, int64_t n_keep, int64_t n_global,
                        int64_t dim, void* stream) {
  MB_REQUIRE(src && out && (n_keep == 0 || ids_keep) && (n_global == 0 || global_tokens),
             "mb_token_gather_fwd: null pointer");
  MB_REQUIRE(dim % 4 == 0, "mb_token_gather_fwd: dim must be a multiple of 4");
  const long long rows = batch * (n_keep + n_global);
  if (rows == 0) return 0;
  token_gather_kernel<<<(unsigned)((rows + 7) / 8), kRowThreads, 0,
                        reinterpret_cast<cudaStream_t>(stream)>>>(
      src, reinterpret_cast<const long long*>(ids_keep), global_tokens, out, (int)batch, (int)n_src,
      (int)n_keep, (int)n_global, (int)dim);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mb_token_gather_bwd(const float* dout, const int64_t* ids_keep, float* dsrc, float* dglobal,
                        int64_t batch, int64_t n_src, int64_t n_keep, int64_t n_global, int64_t dim,
                        void* stream) {
  MB_REQUIRE(dout && dsrc, "mb_token_gather_bwd: null pointer");
  MB_REQUIRE(dim % 4 == 0, "mb_token_gather_bwd: dim must be a multiple of 4");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MB_CHECK_CUDA(cudaMemsetAsync(dsrc, 0, (size_t)batch * n_src * dim * sizeof(float), st));
  const long long rows = batch * n_keep;
  if (rows > 0) {
    token_scatter_kernel<<<(unsigned)((rows + 7) / 8), kRowThreads, 0, st>>>(
        dout, reinterpret_cast<const long long*>(ids_keep), dsrc, (int)batch, (int)n_src,
        (int)n_keep, (int)n_global, (int)dim);
    MB_CHECK_CUDA(cudaGetLastError());
  }
  if (dglobal && n_global > 0) {
    dim3 grid((unsigned)((dim + 255) / 256), (unsigned)n_global);
    global_token_grad_kernel<<<grid, 256, 0, st>>>(dout, dglobal, (int)batch, (int)n_keep,
                                                   (int)n_global, (int)dim);
    MB_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

int mb_fill_global_rows(const float* global_tokens, float* out, int64_t batch, int64_t rows_total,
                        int64_t row_offset, int64_t n_global, int64_t dim, void* stream) {
  MB_REQUIRE(global_tokens && out, "mb_fill_global_rows: null pointer");
  MB_REQUIRE(dim % 4 == 0, "mb_fill_global_rows: dim must be a multiple of 4");
  const long long n = batch * n_global * (dim / 4);
  if (n == 0) return 0;
  fill_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      global_tokens, out, (int)batch, (int)rows_total, (int)row_offset, (int)n_global, (int)dim);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mb_cast_f32_to_bf16(const float* in, void* out, int64_t n, void* stream) {
  MB_REQUIRE(in && out, "mb_cast_f32_to_bf16: null pointer");
  MB_REQUIRE(n % 4 == 0, "mb_cast_f32_to_bf16: n must be a multiple of 4");
  if (n == 0) return 0;
  const long long n4 = n / 4;
  MB_CHECK_CUDA(launch_row(cast_f32_bf16_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0,
                         reinterpret_cast<cudaStream_t>(stream), in, reinterpret_cast<__nv_bfloat16*>(out), n4));
  return 0;
}

}  // extern "C"
