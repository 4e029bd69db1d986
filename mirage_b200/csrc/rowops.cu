// mirage_b200/csrc/rowops.cu
//
// HBM-bound row kernels of the MIRAGE hot path (one warp per token row, 16-byte accesses):
//   LayerNorm forward / backward         nn.LayerNorm(eps=1e-6), mirage/utils.py:260-261
//   visible-token gather / scatter       mirage/model.py:384-391
//   bias gradient (column sums)          backward of every nn.Linear bias
//   fp32 -> bf16 cast, global-token fill
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
                     float eps) {
  const long long row = (long long)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
  if (row >= M) return;
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
// Per-block partial sums of dw = sum dy * xhat and db = sum dy go to part[blk, 2, D]; a second
// kernel reduces them (deterministic, no atomics).
template <int VPL, bool DY_BF16>
__global__ void __launch_bounds__(kRowThreads)
layernorm_bwd_kernel(const void* __restrict__ dy, const float* __restrict__ x,
                     const float* __restrict__ w, const float* __restrict__ mean_in,
                     const float* __restrict__ rstd_in, const float* __restrict__ dres,
                     float* __restrict__ dx, float* __restrict__ part, long long M, int D,
                     long long ldx, long long lddy, long long lddx, int rows_per_block) {
  extern __shared__ float sred[];  // [8 warps][2][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 aw[VPL], ab[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    aw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const long long row0 = (long long)blockIdx.x * rows_per_block;
  for (int rr = warp; rr < rows_per_block; rr += kRowThreads / 32) {
    const long long row = row0 + rr;
    if (row >= M) break;
    const float mean = mean_in[row], rstd = rstd_in[row];
    float4 xh[VPL], g[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int col = (i * 32 + lane) * 4;
      const float4 xv = ld4(x + row * ldx + col);
      float4 d;
      if (DY_BF16) {
        const uint2 pk = *reinterpret_cast<const uint2*>(
            reinterpret_cast<const __nv_bfloat16*>(dy) + row * lddy + col);
        const float2 a = unpack_bf16x2(pk.x), c = unpack_bf16x2(pk.y);
        d = make_float4(a.x, a.y, c.x, c.y);
      } else {
        d = ld4(reinterpret_cast<const float*>(dy) + row * lddy + col);
      }
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + col));
      xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd,
                          (xv.w - mean) * rstd);
      g[i] = make_float4(d.x * wv.x, d.y * wv.y, d.z * wv.z, d.w * wv.w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
      aw[i].x += d.x * xh[i].x; aw[i].y += d.y * xh[i].y;
      aw[i].z += d.z * xh[i].z; aw[i].w += d.w * xh[i].w;
      ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
    }
    const float m1 = warp_sum(s1) / D, m2 = warp_sum(s2) / D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int col = (i * 32 + lane) * 4;
      float4 o;
      o.x = rstd * (g[i].x - m1 - xh[i].x * m2);
      o.y = rstd * (g[i].y - m1 - xh[i].y * m2);
      o.z = rstd * (g[i].z - m1 - xh[i].z * m2);
      o.w = rstd * (g[i].w - m1 - xh[i].w * m2);
      if (dres) {
        const float4 r = ld4(dres + row * lddx + col);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      st4(dx + row * lddx + col, o);
    }
  }
  // block reduction of the parameter-gradient partials
  float* sw = sred + warp * 2 * D;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int col = (i * 32 + lane) * 4;
    st4(sw + col, aw[i]);
    st4(sw + D + col, ab[i]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * D; c += kRowThreads) {
    float acc = 0.f;
#pragma unroll
    for (int wv = 0; wv < kRowThreads / 32; ++wv) acc += sred[wv * 2 * D + c];
    part[(long long)blockIdx.x * 2 * D + c] = acc;
  }
}

// out[c] (+)= sum_r part[r, c]   (c < cols); one thread per column, coalesced across threads
__global__ void colsum_reduce_kernel(const float* __restrict__ part, float* __restrict__ out0,
                                     float* __restrict__ out1, int rows, int D, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 2 * D) return;
  float acc = 0.f;
  for (int r = 0; r < rows; ++r) acc += part[(long long)r * 2 * D + c];
  float* o = (c < D) ? (out0 + c) : (out1 + (c - D));
  *o = accumulate ? (*o + acc) : acc;
}

// ------------------------------------------------------------------------------------------
// column sums of a bf16 / fp32 matrix (bias gradients): part[blk, N] then reduce.
// Each block walks rows_per_block rows; thread t owns 4 consecutive columns per 1024-column slab.
// ------------------------------------------------------------------------------------------
template <bool IN_BF16>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const void* __restrict__ a, float* __restrict__ part, long long M, int N,
                      long long lda, int rows_per_block) {
  const long long row0 = (long long)blockIdx.x * rows_per_block;
  const long long row1 = (row0 + rows_per_block < M) ? row0 + rows_per_block : M;
  for (int col = (blockIdx.y * 256 + threadIdx.x) * 4; col < N; col += gridDim.y * 1024) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long r = row0; r < row1; ++r) {
      if (IN_BF16) {
        const uint2 pk = *reinterpret_cast<const uint2*>(
            reinterpret_cast<const __nv_bfloat16*>(a) + r * lda + col);
        const float2 u = unpack_bf16x2(pk.x), v = unpack_bf16x2(pk.y);
        acc.x += u.x; acc.y += u.y; acc.z += v.x; acc.w += v.y;
      } else {
        const float4 v = ld4(reinterpret_cast<const float*>(a) + r * lda + col);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    st4(part + (long long)blockIdx.x * N + col, acc);
  }
}

__global__ void colsum_final_kernel(const float* __restrict__ part, float* __restrict__ out, int rows,
                                    int N, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float acc = 0.f;
  for (int r = 0; r < rows; ++r) acc += part[(long long)r * N + c];
  out[c] = accumulate ? out[c] + acc : acc;
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
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = ld4(in + i * 4);
  uint2 pk;
  pk.x = pack_bf16x2(v.x, v.y);
  pk.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(out + i * 4) = pk;
}

template <bool OUT_BF16>
static int launch_ln_fwd(const float* x, const float* w, const float* b, void* y, float* mean,
                         float* rstd, long long M, int D, long long ldx, long long ldy, float eps,
                         cudaStream_t st) {
  const unsigned grid = (unsigned)((M + 7) / 8);
#define MB_LN(V)                                                                                   \
  case V:                                                                                          \
    layernorm_fwd_kernel<V, OUT_BF16><<<grid, kRowThreads, 0, st>>>(x, w, b, y, mean, rstd, M, D,  \
                                                                     ldx, ldy, eps);               \
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

int64_t mb_layernorm_bwd_workspace(int64_t rows, int64_t dim) {
  const int64_t blocks = (rows + 63) / 64;
  return blocks * 2 * dim * (int64_t)sizeof(float);
}

int mb_layernorm_bwd(const void* dy, int32_t dy_dtype, const float* x, const float* weight,
                     const float* mean, const float* rstd, const float* dres, float* dx,
                     float* dweight, float* dbias, int32_t accumulate, void* workspace,
                     int64_t rows, int64_t dim, int64_t ldx, int64_t lddy, int64_t lddx,
                     void* stream) {
  MB_REQUIRE(dy && x && weight && mean && rstd && dx && dweight && dbias && workspace,
             "mb_layernorm_bwd: null pointer");
  MB_REQUIRE(dim % 128 == 0 && dim >= 128 && dim <= 1024,
             "mb_layernorm_bwd: dim=%lld unsupported (multiple of 128, <= 1024)", (long long)dim);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rpb = 64;
  const unsigned grid = (unsigned)((rows + rpb - 1) / rpb);
  const size_t smem = (size_t)8 * 2 * dim * sizeof(float);
  float* part = reinterpret_cast<float*>(workspace);
  const int D = (int)dim;
#define MB_LNB(V, BF)                                                                              \
  {                                                                                                \
    auto kern = layernorm_bwd_kernel<V, BF>;                                                       \
    if (smem > 48 * 1024)                                                                          \
      MB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                         (int)smem));                                              \
    kern<<<grid, kRowThreads, smem, st>>>(dy, x, weight, mean, rstd, dres, dx, part, rows, D, ldx, \
                                          lddy, lddx, rpb);                                        \
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
  MB_CHECK_CUDA(cudaGetLastError());
  colsum_reduce_kernel<<<(2 * D + 255) / 256, 256, 0, st>>>(part, dweight, dbias, (int)grid, D,
                                                             accumulate);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int64_t mb_colsum_workspace(int64_t rows, int64_t cols) {
  const int64_t blocks = (rows + 255) / 256;
  return blocks * cols * (int64_t)sizeof(float);
}

int mb_colsum(const void* a, int32_t a_dtype, float* out, int32_t accumulate, void* workspace,
              int64_t rows, int64_t cols, int64_t lda, void* stream) {
  MB_REQUIRE(a && out && workspace, "mb_colsum: null pointer");
  MB_REQUIRE(cols % 4 == 0 && lda % 4 == 0, "mb_colsum: cols and lda must be multiples of 4");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rpb = 256;
  dim3 grid((unsigned)((rows + rpb - 1) / rpb), (unsigned)((cols + 1023) / 1024));
  float* part = reinterpret_cast<float*>(workspace);
  if (a_dtype == MB_BF16)
    colsum_partial_kernel<true><<<grid, 256, 0, st>>>(a, part, rows, (int)cols, lda, rpb);
  else
    colsum_partial_kernel<false><<<grid, 256, 0, st>>>(a, part, rows, (int)cols, lda, rpb);
  MB_CHECK_CUDA(cudaGetLastError());
  colsum_final_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(part, out, (int)grid.x,
                                                                       (int)cols, accumulate);
  MB_CHECK_CUDA(cudaGetLastError());
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
  cast_f32_bf16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0,
                         reinterpret_cast<cudaStream_t>(stream)>>>(
      in, reinterpret_cast<__nv_bfloat16*>(out), n4);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
