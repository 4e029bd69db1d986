// mirage_b200/csrc/pool.cu
//
// Classification tail of the MIRAGE cls wrappers, fused:
//     pooled[b, :] = mean_{t in [row_begin, row_end)} LayerNorm(x[b, t, :])
// mirage_wrapper.py:217-224 (MIRAGEClsGlobal: mean over the patch tokens), :230-233 (CLS: the global
// token), :236-244 (TokenMix: both, concatenated -- two calls with different row ranges / output offsets).
// The reference materialises the normalised [B, N, D] tensor and reduces it with a second kernel; here
// the row statistics, the normalisation and the token mean happen in one pass over x (HBM-bound:
// 4 * D bytes per pooled row), and the backward writes dx in one pass without ever forming dy [B, N, D].
//
//   forward   stage 1: CTA (b, slab) -- one warp per row, x_hat accumulated in registers, per-warp sums
//                      combined through shared memory -> partial[b, slab, D]
//             stage 2: pooled = gamma * (sum_slabs partial) / cnt + beta;  xhat_mean saved for d_gamma
//   backward  dy_t = d_pooled[b] / cnt for every pooled row t:
//             dx_t = rstd_t * (w - mean(w) - xhat_t * mean(w * xhat_t)),  w = gamma * dy;  rows outside the
//             range get zero (optional);  d_gamma = sum_b d_pooled[b] * xhat_mean[b],  d_beta = sum_b d_pooled[b]
#include "../../include/mirage_b200.h"
#include "common.cuh"

namespace mb200 {

constexpr int kPoolThreads = 256;
constexpr int kPoolWarps = kPoolThreads / 32;

__device__ __forceinline__ float4 pld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void pst4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

template <int VPL>
__global__ void __launch_bounds__(kPoolThreads)
ln_meanpool_partial_kernel(const float* __restrict__ x, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                           float* __restrict__ partial, int N, int D, int row_begin, int row_end, int slabs,
                           float eps) {
  extern __shared__ float s_acc[];  // [kPoolWarps][D]
  const int b = blockIdx.x / slabs, slab = blockIdx.x % slabs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cnt = row_end - row_begin;
  const int per = (cnt + slabs - 1) / slabs;
  const int r0 = row_begin + slab * per, r1 = min(row_end, r0 + per);
  float4 acc[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = r0 + warp; r < r1; r += kPoolWarps) {
    const float* xr = x + (static_cast<long long>(b) * N + r) * D;
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      v[i] = pld4(xr + (i * 32 + lane) * 4);
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
      mean_out[static_cast<long long>(b) * N + r] = mean;
      rstd_out[static_cast<long long>(b) * N + r] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      acc[i].x += (v[i].x - mean) * rstd;
      acc[i].y += (v[i].y - mean) * rstd;
      acc[i].z += (v[i].z - mean) * rstd;
      acc[i].w += (v[i].w - mean) * rstd;
    }
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) pst4(s_acc + warp * D + (i * 32 + lane) * 4, acc[i]);
  __syncthreads();
  float* out = partial + (static_cast<long long>(b) * slabs + slab) * D;
  for (int c = threadIdx.x; c < D; c += kPoolThreads) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kPoolWarps; ++w) t += s_acc[w * D + c];
    out[c] = t;
  }
}

__global__ void __launch_bounds__(kPoolThreads)
ln_meanpool_final_kernel(const float* __restrict__ partial, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float* __restrict__ pooled, float* __restrict__ xhat_mean,
                         int B, int D, int slabs, int cnt, long long ld_pooled) {
  const long long idx = static_cast<long long>(blockIdx.x) * kPoolThreads + threadIdx.x;
  if (idx >= static_cast<long long>(B) * D) return;
  const int b = static_cast<int>(idx / D), c = static_cast<int>(idx % D);
  float t = 0.f;
  for (int s = 0; s < slabs; ++s) t += partial[(static_cast<long long>(b) * slabs + s) * D + c];
  const float m = t / cnt;
  xhat_mean[idx] = m;
  pooled[b * ld_pooled + c] = m * gamma[c] + beta[c];
}

template <int VPL>
__global__ void __launch_bounds__(kPoolThreads)
ln_meanpool_bwd_kernel(const float* __restrict__ d_pooled, long long ld_dp, const float* __restrict__ x,
                       const float* __restrict__ gamma, const float* __restrict__ mean_in,
                       const float* __restrict__ rstd_in, float* __restrict__ dx, int B, int N, int D,
                       int row_begin, int row_end, int zero_outside) {
  const long long row = static_cast<long long>(blockIdx.x) * kPoolWarps + (threadIdx.x >> 5);
  if (row >= static_cast<long long>(B) * N) return;
  const int lane = threadIdx.x & 31;
  const int b = static_cast<int>(row / N), t = static_cast<int>(row % N);
  float* dxr = dx + row * D;
  if (t < row_begin || t >= row_end) {
    if (zero_outside) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) pst4(dxr + (i * 32 + lane) * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    return;
  }
  const float inv_cnt = 1.f / (row_end - row_begin), invD = 1.f / D;
  const float mean = mean_in[row], rstd = rstd_in[row];
  float4 xh[VPL], w[VPL];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int col = (i * 32 + lane) * 4;
    const float4 xv = pld4(x + row * D + col);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 dp = __ldg(reinterpret_cast<const float4*>(d_pooled + b * ld_dp + col));
    xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
    w[i] = make_float4(dp.x * inv_cnt * g.x, dp.y * inv_cnt * g.y, dp.z * inv_cnt * g.z, dp.w * inv_cnt * g.w);
    s1 += (w[i].x + w[i].y) + (w[i].z + w[i].w);
    s2 += (w[i].x * xh[i].x + w[i].y * xh[i].y) + (w[i].z * xh[i].z + w[i].w * xh[i].w);
  }
  const float m1 = warp_sum(s1) * invD, m2 = warp_sum(s2) * invD;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float4 o;
    o.x = rstd * (w[i].x - m1 - xh[i].x * m2);
    o.y = rstd * (w[i].y - m1 - xh[i].y * m2);
    o.z = rstd * (w[i].z - m1 - xh[i].z * m2);
    o.w = rstd * (w[i].w - m1 - xh[i].w * m2);
    pst4(dxr + (i * 32 + lane) * 4, o);
  }
}

__global__ void __launch_bounds__(kPoolThreads)
ln_meanpool_param_grad_kernel(const float* __restrict__ d_pooled, long long ld_dp,
                              const float* __restrict__ xhat_mean, float* __restrict__ d_gamma,
                              float* __restrict__ d_beta, int B, int D, int accumulate) {
  const int c = blockIdx.x * kPoolThreads + threadIdx.x;
  if (c >= D) return;
  float dg = 0.f, db = 0.f;
  for (int b = 0; b < B; ++b) {
    const float dp = d_pooled[b * ld_dp + c];
    dg += dp * xhat_mean[static_cast<long long>(b) * D + c];
    db += dp;
  }
  if (accumulate) {
    d_gamma[c] += dg;
    d_beta[c] += db;
  } else {
    d_gamma[c] = dg;
    d_beta[c] = db;
  }
}

static int pool_slabs(int64_t batch, int64_t cnt) {
  // enough CTAs to fill the machine a couple of times, at least 8 rows per slab
  int64_t want = (2 * sm_count() + batch - 1) / batch;
  int64_t max_slabs = (cnt + 7) / 8;
  if (want > max_slabs) want = max_slabs;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  return static_cast<int>(want);
}

}  // namespace mb200

using namespace mb200;

extern "C" {

int64_t mb_ln_meanpool_workspace(int64_t batch, int64_t dim) { return batch * 64 * dim * 4; }

int mb_ln_meanpool_fwd(const float* x, const float* gamma, const float* beta, float* pooled, int64_t ld_pooled,
                       float* xhat_mean, float* mean, float* rstd, void* workspace, int64_t batch, int64_t n_tokens,
                       int64_t dim, int64_t row_begin, int64_t row_end, float eps, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(x && gamma && beta && pooled && xhat_mean && mean && rstd && workspace, "mb_ln_meanpool_fwd: null pointer");
  MB_REQUIRE(dim % 128 == 0 && dim >= 128 && dim <= 1024, "mb_ln_meanpool_fwd: dim %lld unsupported", (long long)dim);
  MB_REQUIRE(0 <= row_begin && row_begin < row_end && row_end <= n_tokens, "mb_ln_meanpool_fwd: bad row range");
  MB_REQUIRE(batch > 0 && batch * n_tokens < (1ll << 31), "mb_ln_meanpool_fwd: batch out of range");
  const int slabs = pool_slabs(batch, row_end - row_begin);
  float* partial = reinterpret_cast<float*>(workspace);
  const size_t smem = static_cast<size_t>(kPoolWarps) * dim * sizeof(float);
  const unsigned grid = static_cast<unsigned>(batch * slabs);
#define MB_POOL_FWD(V)                                                                                   \
  ln_meanpool_partial_kernel<V><<<grid, kPoolThreads, smem, stream>>>(                                   \
      x, mean, rstd, partial, (int)n_tokens, (int)dim, (int)row_begin, (int)row_end, slabs, eps)
  switch (dim / 128) {
    case 1: MB_POOL_FWD(1); break;
    case 2: MB_POOL_FWD(2); break;
    case 3: MB_POOL_FWD(3); break;
    case 4: MB_POOL_FWD(4); break;
    case 5: MB_POOL_FWD(5); break;
    case 6: MB_POOL_FWD(6); break;
    case 7: MB_POOL_FWD(7); break;
    default: MB_POOL_FWD(8); break;
  }
#undef MB_POOL_FWD
  MB_CHECK_CUDA(cudaGetLastError());
  const long long total = batch * dim;
  ln_meanpool_final_kernel<<<static_cast<unsigned>((total + kPoolThreads - 1) / kPoolThreads), kPoolThreads, 0,
                             stream>>>(partial, gamma, beta, pooled, xhat_mean, (int)batch, (int)dim, slabs,
                                       (int)(row_end - row_begin), ld_pooled);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mb_ln_meanpool_bwd(const float* d_pooled, int64_t ld_dp, const float* x, const float* gamma, const float* mean,
                       const float* rstd, const float* xhat_mean, float* dx, float* d_gamma, float* d_beta,
                       int32_t accumulate, int32_t zero_outside, int64_t batch, int64_t n_tokens, int64_t dim,
                       int64_t row_begin, int64_t row_end, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(d_pooled && x && gamma && mean && rstd && xhat_mean && dx && d_gamma && d_beta,
             "mb_ln_meanpool_bwd: null pointer");
  MB_REQUIRE(dim % 128 == 0 && dim >= 128 && dim <= 1024, "mb_ln_meanpool_bwd: dim %lld unsupported", (long long)dim);
  MB_REQUIRE(0 <= row_begin && row_begin < row_end && row_end <= n_tokens, "mb_ln_meanpool_bwd: bad row range");
  MB_REQUIRE(ld_dp % 4 == 0, "mb_ln_meanpool_bwd: ld_dp must be a multiple of 4");
  const long long rows = batch * n_tokens;
  const unsigned grid = static_cast<unsigned>((rows + kPoolWarps - 1) / kPoolWarps);
#define MB_POOL_BWD(V)                                                                                    \
  ln_meanpool_bwd_kernel<V><<<grid, kPoolThreads, 0, stream>>>(d_pooled, ld_dp, x, gamma, mean, rstd, dx, \
                                                               (int)batch, (int)n_tokens, (int)dim,       \
                                                               (int)row_begin, (int)row_end, zero_outside)
  switch (dim / 128) {
    case 1: MB_POOL_BWD(1); break;
    case 2: MB_POOL_BWD(2); break;
    case 3: MB_POOL_BWD(3); break;
    case 4: MB_POOL_BWD(4); break;
    case 5: MB_POOL_BWD(5); break;
    case 6: MB_POOL_BWD(6); break;
    case 7: MB_POOL_BWD(7); break;
    default: MB_POOL_BWD(8); break;
  }
#undef MB_POOL_BWD
  MB_CHECK_CUDA(cudaGetLastError());
  ln_meanpool_param_grad_kernel<<<static_cast<unsigned>((dim + kPoolThreads - 1) / kPoolThreads), kPoolThreads, 0,
                                  stream>>>(d_pooled, ld_dp, xhat_mean, d_gamma, d_beta, (int)batch, (int)dim,
                                            accumulate);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
