// mirage_b200/csrc/optim.cu
//
// The optimizer step either side of the MIRAGE hot path:
//   AdamW over parameter groups   mutils/optim_factory.py:33-92 (groups), :171-172 (optim.AdamW)
//   per-step lr / wd assignment   run_pretraining.py:683-688
//   grad-norm / clip / skip       mutils/native_scaler.py:16-37, :46-61
// as ONE pass over the gradient buckets instead of torch's multi-tensor AdamW + a per-parameter norm
// + one fp32->bf16 cast kernel per weight matrix:
//
//   mb_optim_prepare   1 block : (clip / skip mode) global norm from the partials of mb_sumsq, clip
//                                coefficient, skip flag; step counter + bias corrections
//   mb_adamw_step      N blocks: p, m, v updated in fp32 (torch.optim.AdamW arithmetic), the bf16 weight
//                                shadow the GEMMs read refreshed in place, the gradient zeroed for the
//                                next step, sum(g^2) partials of the (unclipped) gradient
//   mb_optim_finish    1 block : global gradient norm from those partials (when not clipping)
//
// Every parameter is a SEGMENT {p, g, m, v, shadow, numel, group}; a block owns kChunk consecutive
// elements of one segment (block -> segment through a prefix table, binary search).  Hyper-parameters live in
// a small device array indexed by group, so a step captured in a CUDA graph follows the host's lr / wd
// schedule without re-capture.  HBM-bound: 4 (g) + 3 x 8 (p, m, v read + write) + 4 (g zeroed) + 2 (shadow)
// = 34 bytes per parameter with shadow and zeroing, 28 for the bare update.
#include "../../include/mirage_b200.h"
#include "common.cuh"

namespace mb200 {

constexpr int kOptThreads = 256;
constexpr int kOptChunk = 4096;  // elements per block: 4 x float4 per thread

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < kOptThreads / 32) ? red[threadIdx.x] : 0.f;
  if (warp == 0) t = warp_sum(t);
  return t;  // valid in warp 0
}

// sum of squares of a flat fp32 buffer -> one partial per block (deterministic two-stage reduction)
__global__ void __launch_bounds__(kOptThreads) sumsq_kernel(const float* __restrict__ x, long long n,
                                                            float* __restrict__ partials) {
  __shared__ float red[kOptThreads / 32];
  float acc = 0.f;
  const long long n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = (long long)blockIdx.x * kOptThreads + threadIdx.x; i < n4;
       i += (long long)gridDim.x * kOptThreads) {
    const float4 v = x4[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += kOptThreads) acc += x[i] * x[i];
  const float t = block_sum_256(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kOptThreads)
optim_prepare_kernel(mb_optim_state* st, const float* partials, int n_partials, float clip_norm,
                     float skip_norm, float beta1, float beta2) {
  __shared__ float red[kOptThreads / 32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n_partials; i += kOptThreads) acc += partials[i];
  const float t = block_sum_256(acc, red);
  if (threadIdx.x == 0) {
    float coef = 1.f;
    int skip = 0;
    if (n_partials > 0) {
      const float norm = sqrtf(t);
      st->grad_norm = norm;
      // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
      if (clip_norm > 0.f) coef = fminf(1.f, clip_norm / (norm + 1e-6f));
      // native_scaler.py:28-32: a step whose norm reaches skip_grad is dropped
      if (skip_norm > 0.f && !(norm < skip_norm)) skip = 1;
    }
    st->clip_coef = coef;
    st->skipped = skip;
    if (!skip) {
      const long long step = st->step + 1;
      st->step = step;
      const double s = static_cast<double>(step);
      st->bias_corr1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), s));
      st->bias_corr2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), s)));
    }
  }
}

__global__ void __launch_bounds__(kOptThreads)
optim_finish_kernel(mb_optim_state* st, const float* partials, int n_partials) {
  __shared__ float red[kOptThreads / 32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n_partials; i += kOptThreads) acc += partials[i];
  const float t = block_sum_256(acc, red);
  if (threadIdx.x == 0) st->grad_norm = sqrtf(t);
}

struct AdamElem {
  float p, m, v;
};

__device__ __forceinline__ AdamElem adam_one(float p, float g, float m, float v, float lr, float wd,
                                             float beta1, float beta2, float eps, float bc1, float bc2s) {
  // torch.optim.AdamW (single-tensor form): decoupled decay, then the bias-corrected Adam update
  p = p * (1.f - lr * wd);
  m = m + (g - m) * (1.f - beta1);           // exp_avg.lerp_(grad, 1 - beta1)
  v = v * beta2 + (1.f - beta2) * g * g;     // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / bc2s + eps;
  p = p - (lr / bc1) * (m / denom);
  return {p, m, v};
}

__global__ void __launch_bounds__(kOptThreads)
adamw_kernel(const mb_optim_segment* __restrict__ segs, const int* __restrict__ block_prefix, int n_segs,
             const mb_optim_hyper* __restrict__ hyper, const mb_optim_state* __restrict__ st,
             float beta1, float beta2, float eps, int zero_grad, float* __restrict__ partials) {
  __shared__ float red[kOptThreads / 32];
  __shared__ int s_seg;
  if (threadIdx.x == 0) {
    // block_prefix[i] = first block of segment i; block_prefix[n_segs] = total
    int lo = 0, hi = n_segs - 1;
    const int b = blockIdx.x;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (block_prefix[mid] <= b) lo = mid; else hi = mid - 1;
    }
    s_seg = lo;
  }
  __syncthreads();
  const mb_optim_segment sg = segs[s_seg];
  const long long base = static_cast<long long>(blockIdx.x - block_prefix[s_seg]) * kOptChunk;
  const mb_optim_hyper hp = hyper[sg.group];
  const float lr = hp.lr * hp.lr_scale, wd = hp.weight_decay;
  const float coef = st->clip_coef, bc1 = st->bias_corr1, bc2s = st->bias_corr2_sqrt;
  const bool skip = st->skipped != 0;
  float* __restrict__ P = sg.param;
  float* __restrict__ G = sg.grad;
  float* __restrict__ M = sg.exp_avg;
  float* __restrict__ V = sg.exp_avg_sq;
  __nv_bfloat16* __restrict__ S = reinterpret_cast<__nv_bfloat16*>(sg.shadow);
  float acc = 0.f;
  const long long n = sg.numel;
  const bool vec = (sg.flags & 1) != 0;  // every pointer 16-byte aligned (8 for the shadow)
  if (vec) {
#pragma unroll
    for (int it = 0; it < kOptChunk / (kOptThreads * 4); ++it) {
      const long long i = base + (static_cast<long long>(it) * kOptThreads + threadIdx.x) * 4;
      if (i + 3 < n) {
        const float4 g4 = *reinterpret_cast<const float4*>(G + i);
        acc += g4.x * g4.x + g4.y * g4.y + g4.z * g4.z + g4.w * g4.w;
        if (!skip) {
          const float4 p4 = *reinterpret_cast<const float4*>(P + i);
          const float4 m4 = *reinterpret_cast<const float4*>(M + i);
          const float4 v4 = *reinterpret_cast<const float4*>(V + i);
          const AdamElem a = adam_one(p4.x, g4.x * coef, m4.x, v4.x, lr, wd, beta1, beta2, eps, bc1, bc2s);
          const AdamElem b = adam_one(p4.y, g4.y * coef, m4.y, v4.y, lr, wd, beta1, beta2, eps, bc1, bc2s);
          const AdamElem c = adam_one(p4.z, g4.z * coef, m4.z, v4.z, lr, wd, beta1, beta2, eps, bc1, bc2s);
          const AdamElem d = adam_one(p4.w, g4.w * coef, m4.w, v4.w, lr, wd, beta1, beta2, eps, bc1, bc2s);
          *reinterpret_cast<float4*>(P + i) = make_float4(a.p, b.p, c.p, d.p);
          *reinterpret_cast<float4*>(M + i) = make_float4(a.m, b.m, c.m, d.m);
          *reinterpret_cast<float4*>(V + i) = make_float4(a.v, b.v, c.v, d.v);
          if (S != nullptr) {
            uint2 pk;
            pk.x = pack_bf16x2(a.p, b.p);
            pk.y = pack_bf16x2(c.p, d.p);
            *reinterpret_cast<uint2*>(S + i) = pk;
          }
        }
        if (zero_grad) *reinterpret_cast<float4*>(G + i) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        for (long long j = i; j < n && j < i + 4; ++j) {
          const float g = G[j];
          acc += g * g;
          if (!skip) {
            const AdamElem a = adam_one(P[j], g * coef, M[j], V[j], lr, wd, beta1, beta2, eps, bc1, bc2s);
            P[j] = a.p; M[j] = a.m; V[j] = a.v;
            if (S != nullptr) S[j] = __float2bfloat16(a.p);
          }
          if (zero_grad) G[j] = 0.f;
        }
      }
    }
  } else {
    for (int it = 0; it < kOptChunk / kOptThreads; ++it) {
      const long long j = base + static_cast<long long>(it) * kOptThreads + threadIdx.x;
      if (j < n) {
        const float g = G[j];
        acc += g * g;
        if (!skip) {
          const AdamElem a = adam_one(P[j], g * coef, M[j], V[j], lr, wd, beta1, beta2, eps, bc1, bc2s);
          P[j] = a.p; M[j] = a.m; V[j] = a.v;
          if (S != nullptr) S[j] = __float2bfloat16(a.p);
        }
        if (zero_grad) G[j] = 0.f;
      }
    }
  }
  if (partials != nullptr) {
    const float t = block_sum_256(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
  }
}

}  // namespace mb200

using namespace mb200;

extern "C" {

int64_t mb_optim_blocks(int64_t numel) { return (numel + kOptChunk - 1) / kOptChunk; }

int mb_sumsq(const float* x, int64_t n, float* partials, int32_t n_partials, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(x != nullptr && partials != nullptr && n >= 0 && n_partials > 0, "mb_sumsq: bad arguments");
  MB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "mb_sumsq: buffer must be 16-byte aligned");
  sumsq_kernel<<<static_cast<unsigned>(n_partials), kOptThreads, 0, stream>>>(x, n, partials);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mb_optim_prepare(mb_optim_state* state, const float* partials, int32_t n_partials, float clip_norm,
                     float skip_norm, float beta1, float beta2, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(state != nullptr, "mb_optim_prepare: null state");
  MB_REQUIRE(n_partials == 0 || partials != nullptr, "mb_optim_prepare: partials missing");
  optim_prepare_kernel<<<1, kOptThreads, 0, stream>>>(state, partials, n_partials, clip_norm, skip_norm, beta1,
                                                      beta2);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mb_adamw_step(const mb_optim_segment* segments, const int32_t* block_prefix, int32_t n_segments,
                  int64_t n_blocks, const mb_optim_hyper* hyper, const mb_optim_state* state, float beta1,
                  float beta2, float eps, int32_t zero_grad, float* partials, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(segments && block_prefix && hyper && state, "mb_adamw_step: null table pointer");
  MB_REQUIRE(n_segments > 0 && n_blocks > 0 && n_blocks < (1ll << 31), "mb_adamw_step: %d segments, %lld blocks",
             n_segments, (long long)n_blocks);
  adamw_kernel<<<static_cast<unsigned>(n_blocks), kOptThreads, 0, stream>>>(
      segments, block_prefix, n_segments, hyper, state, beta1, beta2, eps, zero_grad, partials);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mb_optim_finish(mb_optim_state* state, const float* partials, int64_t n_partials, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(state && partials && n_partials > 0 && n_partials < (1ll << 31), "mb_optim_finish: bad arguments");
  optim_finish_kernel<<<1, kOptThreads, 0, stream>>>(state, partials, static_cast<int>(n_partials));
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
