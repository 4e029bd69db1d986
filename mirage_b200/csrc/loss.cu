// mirage_b200/csrc/loss.cu
//
// Masked reconstruction criteria of MultiMAE pretraining, fused (HBM-bound: prediction and target are
// each read exactly once per pass):
//   MaskedMSELoss            mirage/criterion.py:87-117   (norm_pix = False)
//   MaskedCrossEntropyLoss   mirage/criterion.py:31-51    (optional label smoothing)
//
//   per_b = sum_pix mask_up[b,pix] * e[b,pix] / sum_pix mask_up[b,pix]      e = mean_c (p-t)^2  |  CE
//   loss  = nanmean_b(per_b)      (samples whose mask is empty or whose loss is NaN are skipped; 0 if every mask
//                                  is empty; without a mask the plain mean, where a NaN propagates)
// CE labels: -100 (torch's default ignore_index) gives zero loss and zero gradient but stays in the denominator,
// as in the reference; any other label outside [0, C) makes the loss +inf (torch would raise a device assert).
// One deviation: a NaN prediction under a ZERO mask pixel is skipped here, whereas the reference's `loss * mask`
// turns it into NaN and drops the whole sample.
//
// Pass 1 writes per-(sample, chunk) partial sums (deterministic, no atomics); a one-block finalize
// kernel produces the scalar loss and, for the backward pass, coef[b] = 1 / (masked_pixels_b *
// n_valid_samples) (0 for skipped samples).  The host never reads a value back (the reference's
// `mask.sum() == 0` host sync, criterion.py:36/:103, is evaluated on the device).
#include "../../include/mirage_b200.h"
#include "common.cuh"

namespace mb200 {

constexpr int kLossThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* s_red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < kLossThreads / 32) ? s_red[threadIdx.x] : 0.f;
  if (warp == 0) t = warp_sum(t);
  __syncthreads();
  return t;  // valid in warp 0
}

// mask value (0/1) of the patch containing pixel (h, w); mask == nullptr means "everything masked"
__device__ __forceinline__ float mask_at(const long long* mask, int b, int h, int w, int scale, int nw,
                                         int n_tok) {
  if (mask == nullptr) return 1.f;
  return (float)mask[(long long)b * n_tok + (h / scale) * nw + (w / scale)];
}

__global__ void __launch_bounds__(kLossThreads)
mse_partial_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                   const long long* __restrict__ mask, float* __restrict__ part, int C, int H, int W,
                   int scale) {
  __shared__ float s_red[kLossThreads / 32];
  const int b = blockIdx.y;
  const int nw = W / scale, n_tok = (H / scale) * nw;
  const long long hw = (long long)H * W;
  const long long n4 = hw / 4;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * kLossThreads + threadIdx.x; i < n4;
       i += (long long)gridDim.x * kLossThreads) {
    const long long pix = i * 4;
    const int h = (int)(pix / W), w = (int)(pix % W);
    const float mk = mask_at(mask, b, h, w, scale, nw, n_tok);  // 4 pixels share a patch (scale % 4 == 0)
    if (mk == 0.f) continue;
    float e = 0.f;
    for (int c = 0; c < C; ++c) {
      const long long off = ((long long)b * C + c) * hw + pix;
      const float4 p = *reinterpret_cast<const float4*>(pred + off);
      const float4 t = *reinterpret_cast<const float4*>(tgt + off);
      const float d0 = p.x - t.x, d1 = p.y - t.y, d2 = p.z - t.z, d3 = p.w - t.w;
      e += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    acc += mk * e;
  }
  const float tot = block_sum(acc, s_red);
  if (threadIdx.x == 0) part[(long long)b * gridDim.x + blockIdx.x] = tot / C;
}

__global__ void __launch_bounds__(kLossThreads)
mse_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
               const long long* __restrict__ mask, const float* __restrict__ coef,
               const float* __restrict__ gout, float* __restrict__ dpred, int C, int H, int W,
               int scale) {
  const int b = blockIdx.y;
  const int nw = W / scale, n_tok = (H / scale) * nw;
  const long long hw = (long long)H * W;
  const long long n4 = hw / 4;
  const float k = gout[0] * coef[b] * 2.f / C;
  for (long long i = (long long)blockIdx.x * kLossThreads + threadIdx.x; i < n4;
       i += (long long)gridDim.x * kLossThreads) {
    const long long pix = i * 4;
    const int h = (int)(pix / W), w = (int)(pix % W);
    const float mk = mask_at(mask, b, h, w, scale, nw, n_tok) * k;
    for (int c = 0; c < C; ++c) {
      const long long off = ((long long)b * C + c) * hw + pix;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mk != 0.f) {
        const float4 p = *reinterpret_cast<const float4*>(pred + off);
        const float4 t = *reinterpret_cast<const float4*>(tgt + off);
        o = make_float4(mk * (p.x - t.x), mk * (p.y - t.y), mk * (p.z - t.z), mk * (p.w - t.w));
      }
      *reinterpret_cast<float4*>(dpred + off) = o;
    }
  }
}

constexpr int kCeMaxRegC = 16;          // class counts up to this are held in registers (13 retinal layers)
constexpr long long kIgnoreIndex = -100;  // torch.nn.functional.cross_entropy default (criterion.py:33 passes none)

// one thread per pixel; channel values are strided by H*W (coalesced across the warp per channel)
__global__ void __launch_bounds__(kLossThreads)
ce_partial_kernel(const float* __restrict__ logits, const long long* __restrict__ tgt,
                  const long long* __restrict__ mask, float* __restrict__ part, int C, int H, int W,
                  int scale, float smoothing) {
  __shared__ float s_red[kLossThreads / 32];
  const int b = blockIdx.y;
  const int nw = W / scale, n_tok = (H / scale) * nw;
  const long long hw = (long long)H * W;
  float acc = 0.f;
  for (long long pix = (long long)blockIdx.x * kLossThreads + threadIdx.x; pix < hw;
       pix += (long long)gridDim.x * kLossThreads) {
    const int h = (int)(pix / W), w = (int)(pix % W);
    const float mk = mask_at(mask, b, h, w, scale, nw, n_tok);
    if (mk == 0.f) continue;
    const float* l = logits + (long long)b * C * hw + pix;
    const long long t = tgt[(long long)b * hw + pix];
    if (t == kIgnoreIndex) continue;  // F.cross_entropy's default ignore_index: zero loss, still in the denominator
    if (t < 0 || t >= C) {            // torch raises a device assert; here the loss goes to +inf (visible, no sync)
      acc = INFINITY;
      continue;
    }
    float mx = -INFINITY, se = 0.f, sl = 0.f, lt = 0.f;
    if (C <= kCeMaxRegC) {
      // all channel loads in flight at once (the one-load-per-iteration online form below was latency-bound:
      // 130 us for the 218 MB of cfg-4 logits), then max / sum-exp over registers
      float v[kCeMaxRegC];
#pragma unroll
      for (int c = 0; c < kCeMaxRegC; ++c) v[c] = c < C ? l[(long long)c * hw] : -INFINITY;
#pragma unroll
      for (int c = 0; c < kCeMaxRegC; ++c) mx = fmaxf(mx, v[c]);
#pragma unroll
      for (int c = 0; c < kCeMaxRegC; ++c) {
        if (c < C) {
          se += __expf(v[c] - mx);
          sl += v[c];
          if (c == t) lt = v[c];
        }
      }
    } else {
      for (int c = 0; c < C; ++c) {
        const float v = l[(long long)c * hw];
        const float nm = fmaxf(mx, v);
        se = se * __expf(mx - nm) + __expf(v - nm);
        mx = nm;
        sl += v;
        if (c == t) lt = v;
      }
    }
    const float lse = mx + __logf(se);
    const float nll = lse - lt;
    acc += mk * ((1.f - smoothing) * nll + smoothing * (lse - sl / C));
  }
  const float tot = block_sum(acc, s_red);
  if (threadIdx.x == 0) part[(long long)b * gridDim.x + blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kLossThreads)
ce_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ tgt,
              const long long* __restrict__ mask, const float* __restrict__ coef,
              const float* __restrict__ gout, float* __restrict__ dlogits, int C, int H, int W,
              int scale, float smoothing) {
  const int b = blockIdx.y;
  const int nw = W / scale, n_tok = (H / scale) * nw;
  const long long hw = (long long)H * W;
  const float k = gout[0] * coef[b];
  for (long long pix = (long long)blockIdx.x * kLossThreads + threadIdx.x; pix < hw;
       pix += (long long)gridDim.x * kLossThreads) {
    const int h = (int)(pix / W), w = (int)(pix % W);
    const float mk = mask_at(mask, b, h, w, scale, nw, n_tok) * k;
    const float* l = logits + (long long)b * C * hw + pix;
    float* d = dlogits + (long long)b * C * hw + pix;
    if (mk == 0.f) {
      for (int c = 0; c < C; ++c) d[(long long)c * hw] = 0.f;
      continue;
    }
    const long long t = tgt[(long long)b * hw + pix];
    if (t < 0 || t >= C) {  // ignore_index (and invalid labels, whose forward loss is already +inf): no gradient
      for (int c = 0; c < C; ++c) d[(long long)c * hw] = 0.f;
      continue;
    }
    float mx = -INFINITY, se = 0.f;
    if (C <= kCeMaxRegC) {
      float v[kCeMaxRegC];
#pragma unroll
      for (int c = 0; c < kCeMaxRegC; ++c) v[c] = c < C ? l[(long long)c * hw] : -INFINITY;
#pragma unroll
      for (int c = 0; c < kCeMaxRegC; ++c) mx = fmaxf(mx, v[c]);
#pragma unroll
      for (int c = 0; c < kCeMaxRegC; ++c) {
        v[c] = __expf(v[c] - mx);   // exp(-inf) = 0 for the padding
        se += v[c];
      }
      const float inv = 1.f / se;
#pragma unroll
      for (int c = 0; c < kCeMaxRegC; ++c) {
        if (c < C) {
          const float hot = (c == t) ? (1.f - smoothing) : 0.f;
          d[(long long)c * hw] = mk * (v[c] * inv - hot - smoothing / C);
        }
      }
      continue;
    }
    for (int c = 0; c < C; ++c) {
      const float v = l[(long long)c * hw];
      const float nm = fmaxf(mx, v);
      se = se * __expf(mx - nm) + __expf(v - nm);
      mx = nm;
    }
    const float inv = 1.f / se;
    for (int c = 0; c < C; ++c) {
      const float sm = __expf(l[(long long)c * hw] - mx) * inv;
      const float hot = (c == t) ? (1.f - smoothing) : 0.f;
      d[(long long)c * hw] = mk * (sm - hot - smoothing / C);
    }
  }
}

// single block: reduce partials, apply the reference's per-sample normalisation and nanmean
__global__ void loss_finalize_kernel(const float* __restrict__ part, const long long* __restrict__ mask,
                                     float* __restrict__ loss, float* __restrict__ coef, int B,
                                     int chunks, int n_tok, int scale, long long pixels) {
  extern __shared__ float s_per[];  // [B] numerators, then [B] denominators
  float* s_num = s_per;
  float* s_den = s_per + B;
  // one warp per sample: lanes stride over the chunk partials and the mask tokens (coalesced), fixed-order
  // shuffle reduction -> deterministic
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int b = warp; b < B; b += nwarps) {
    float n = 0.f;
    for (int c = lane; c < chunks; c += 32) n += part[(long long)b * chunks + c];
    n = warp_sum(n);
    float d;
    if (mask == nullptr) {
      d = (float)pixels;
    } else {
      float cnt = 0.f;  // token counts are far below 2^24: exact in fp32
      for (int t = lane; t < n_tok; t += 32) cnt += (float)mask[(long long)b * n_tok + t];
      d = warp_sum(cnt) * scale * scale;
    }
    if (lane == 0) {
      s_num[b] = n;
      s_den[b] = d;
    }
  }
  __syncthreads();
  // masked form: nanmean over the per-sample means (criterion.py:49, :107) -- samples with an empty mask (0/0) or a
  // NaN loss are skipped; unmasked form: plain mean, a NaN propagates ("we want it to stop training", :51).
  // Block-wide tree reduction in a fixed order (a serial loop over the batch in one thread cost 30 us).
  auto counts = [&](int b) { return s_den[b] > 0.f && (mask == nullptr || s_num[b] == s_num[b]); };
  __shared__ float s_wtot[32], s_wcnt[32];
  float t = 0.f, n = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x)
    if (counts(b)) {
      n += 1.f;
      t += s_num[b] / s_den[b];
    }
  t = warp_sum(t);
  n = warp_sum(n);
  if (lane == 0) {
    s_wtot[warp] = t;
    s_wcnt[warp] = n;
  }
  __syncthreads();
  if (warp == 0) {
    t = lane < nwarps ? s_wtot[lane] : 0.f;
    n = lane < nwarps ? s_wcnt[lane] : 0.f;
    t = warp_sum(t);
    n = warp_sum(n);
    if (lane == 0) {
      s_wtot[0] = t;
      s_wcnt[0] = n;
      loss[0] = n > 0.f ? t / n : 0.f;
    }
  }
  __syncthreads();
  const float valid = s_wcnt[0];
  for (int b = threadIdx.x; b < B; b += blockDim.x)
    coef[b] = (counts(b) && valid > 0.f) ? 1.f / (s_den[b] * valid) : 0.f;
}

constexpr int kMaxLossChunks = 64;

static int loss_chunks(long long hw) {
  long long c = (hw / 4 + kLossThreads * 8 - 1) / (kLossThreads * 8);
  return (int)(c < 1 ? 1 : (c > kMaxLossChunks ? kMaxLossChunks : c));
}

// CE handles ONE pixel (all channels) per thread and iteration: two pixels per thread keep enough loads in flight
// (with the MSE chunking -- 32 pixels per thread at 128 x 128 -- the pass ran at 1.5 TB/s)
static int ce_chunks(long long hw) {
  long long c = (hw + kLossThreads * 2 - 1) / (kLossThreads * 2);
  return (int)(c < 1 ? 1 : (c > kMaxLossChunks ? kMaxLossChunks : c));
}

}  // namespace mb200

using namespace mb200;

extern "C" {

int64_t mb_masked_loss_workspace(int64_t batch, int64_t height, int64_t width) {
  (void)height;
  (void)width;
  return batch * kMaxLossChunks * (int64_t)sizeof(float);   // partial sums, [batch][chunks <= 64]
}

int mb_masked_mse_fwd(const float* pred, const float* target, const int64_t* mask, float* loss,
                      float* coef, void* workspace, int64_t batch, int64_t channels, int64_t height,
                      int64_t width, int32_t scale, void* stream) {
  MB_REQUIRE(pred && target && loss && coef && workspace, "mb_masked_mse_fwd: null pointer");
  MB_REQUIRE(scale > 0 && scale % 4 == 0 && height % scale == 0 && width % scale == 0,
             "mb_masked_mse_fwd: scale %d must be a multiple of 4 dividing %lldx%lld", scale,
             (long long)height, (long long)width);
  MB_REQUIRE(batch > 0 && batch <= 4096, "mb_masked_mse_fwd: batch %lld out of range", (long long)batch);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int chunks = loss_chunks(height * width);
  float* part = reinterpret_cast<float*>(workspace);
  mse_partial_kernel<<<dim3(chunks, (unsigned)batch), kLossThreads, 0, st>>>(
      pred, target, reinterpret_cast<const long long*>(mask), part, (int)channels, (int)height,
      (int)width, scale);
  MB_CHECK_CUDA(cudaGetLastError());
  loss_finalize_kernel<<<1, 1024, 2 * batch * sizeof(float), st>>>(
      part, reinterpret_cast<const long long*>(mask), loss, coef, (int)batch, chunks,
      (int)((height / scale) * (width / scale)), scale, height * width);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mb_masked_mse_bwd(const float* pred, const float* target, const int64_t* mask, const float* coef,
                      const float* grad_out, float* dpred, int64_t batch, int64_t channels,
                      int64_t height, int64_t width, int32_t scale, void* stream) {
  MB_REQUIRE(pred && target && coef && grad_out && dpred, "mb_masked_mse_bwd: null pointer");
  const int chunks = loss_chunks(height * width);
  mse_bwd_kernel<<<dim3(chunks, (unsigned)batch), kLossThreads, 0,
                   reinterpret_cast<cudaStream_t>(stream)>>>(
      pred, target, reinterpret_cast<const long long*>(mask), coef, grad_out, dpred, (int)channels,
      (int)height, (int)width, scale);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mb_masked_ce_fwd(const float* logits, const int64_t* target, const int64_t* mask, float* loss,
                     float* coef, void* workspace, int64_t batch, int64_t channels, int64_t height,
                     int64_t width, int32_t scale, float label_smoothing, void* stream) {
  MB_REQUIRE(logits && target && loss && coef && workspace, "mb_masked_ce_fwd: null pointer");
  MB_REQUIRE(scale > 0 && height % scale == 0 && width % scale == 0,
             "mb_masked_ce_fwd: scale %d must divide %lldx%lld", scale, (long long)height,
             (long long)width);
  MB_REQUIRE(batch > 0 && batch <= 4096, "mb_masked_ce_fwd: batch %lld out of range", (long long)batch);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int chunks = ce_chunks(height * width);
  float* part = reinterpret_cast<float*>(workspace);
  ce_partial_kernel<<<dim3(chunks, (unsigned)batch), kLossThreads, 0, st>>>(
      logits, reinterpret_cast<const long long*>(target), reinterpret_cast<const long long*>(mask),
      part, (int)channels, (int)height, (int)width, scale, label_smoothing);
  MB_CHECK_CUDA(cudaGetLastError());
  loss_finalize_kernel<<<1, 1024, 2 * batch * sizeof(float), st>>>(
      part, reinterpret_cast<const long long*>(mask), loss, coef, (int)batch, chunks,
      (int)((height / scale) * (width / scale)), scale, height * width);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mb_masked_ce_bwd(const float* logits, const int64_t* target, const int64_t* mask,
                     const float* coef, const float* grad_out, float* dlogits, int64_t batch,
                     int64_t channels, int64_t height, int64_t width, int32_t scale,
                     float label_smoothing, void* stream) {
  MB_REQUIRE(logits && target && coef && grad_out && dlogits, "mb_masked_ce_bwd: null pointer");
  const int chunks = ce_chunks(height * width);
  ce_bwd_kernel<<<dim3(chunks, (unsigned)batch), kLossThreads, 0,
                  reinterpret_cast<cudaStream_t>(stream)>>>(
      logits, reinterpret_cast<const long long*>(target), reinterpret_cast<const long long*>(mask),
      coef, grad_out, dlogits, (int)channels, (int)height, (int)width, scale, label_smoothing);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
