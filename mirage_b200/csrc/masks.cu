// mirage_b200/csrc/masks.cu
//
// On-device MultiMAE token-mask sampling: ONE kernel, one CTA per sample, instead
// of the reference's CPU Dirichlet draw + host->device copy + ~20 torch launches and four argsorts per
// step (mirage/model.py:168-239, sample_alphas :145-166).  Same distribution, different (documented)
// random stream -- the reference path stays available bit-exactly in mirage_b200/model.py.
//
// Per sample b (mirage/model.py line numbers in brackets):
//   1. shares ~ Dirichlet(alpha_t)  [:203-207]   Gamma(alpha_t) by Marsaglia-Tsang in log space, normalised;
//      with uniform_tasks the concentration of a uniformly chosen non-empty task subset is alpha_t + 1e-5
//      and 1e-5 for the others  [:145-166]
//   2. per_task_t = round_half_even(share_t * n_encoded)  [:209]
//   3. per task: uniform noise, rank of every token inside its task, mask = rank >= per_task  [:214-222]
//   4. global order: sort key (mask, fresh noise) -> visible tokens first, each class in random order;
//      ids_restore = rank, ids_keep = the first n_encoded of the order  [:224-227]
//   5. final binary mask from the order (the rounding of step 2 need not sum to n_encoded)  [:230-237]
//
// Random stream (Philox4x32-10, counter-based, so results do not depend on the launch geometry):
//   key     = (seed_lo, seed_hi)
//   counter = (draw_lo, draw_hi, b, w)        draw = *draw_counter at launch (the kernel advances it by 1)
//     w = i            (token i):   word 0 = per-task noise, word 1 = global-order noise
//     w = 2^30 + 64 t + a  (task t, attempt a): words 0,1 = Box-Muller normal, 2 = acceptance uniform,
//                                               3 = boost uniform (alpha < 1)
//     w = 2^31         : word 0 = task-subset choice (uniform_tasks)
// Ranks are computed on the raw 32-bit words (ties broken by token index), i.e. exact uniform
// permutations up to 2^-32 tie probability.
#include "../../include/mirage_b200.h"
#include "common.cuh"

namespace mb200 {

constexpr int kMaskThreads = 256;
constexpr int kMaskMaxTasks = 8;

struct Philox4 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                                 uint32_t c2, uint32_t c3) {
  constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(kM0, c0), lo0 = kM0 * c0;
    const uint32_t hi1 = __umulhi(kM1, c2), lo1 = kM1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += kW0; k1 += kW1;
  }
  return {c0, c1, c2, c3};
}

__device__ __forceinline__ double u01_open(uint32_t x) {  // (0, 1)
  return (static_cast<double>(x) + 0.5) * (1.0 / 4294967296.0);
}

struct MaskParams {
  int n_tasks;
  int counts[kMaskMaxTasks];
  float alphas[kMaskMaxTasks];
  int n_all, n_enc, uniform_tasks;
};

// log of a Gamma(alpha, 1) variate (log space: alpha = 1e-5 under uniform_tasks underflows otherwise)
__device__ double log_gamma_variate(double alpha, uint32_t k0, uint32_t k1, uint32_t d0, uint32_t d1,
                                    uint32_t b, int t) {
  const double a = alpha < 1.0 ? alpha + 1.0 : alpha;
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  double boost_u = 0.5;
  double g = d;
  for (int attempt = 0; attempt < 64; ++attempt) {
    const Philox4 r = philox4x32_10(k0, k1, d0, d1, b, (1u << 30) + 64u * t + attempt);
    if (attempt == 0) boost_u = u01_open(r.w);
    const double u1 = u01_open(r.x), u2 = u01_open(r.y), u = u01_open(r.z);
    const double nrm = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    const double v0 = 1.0 + c * nrm;
    if (v0 <= 0.0) continue;
    const double v = v0 * v0 * v0;
    if (log(u) < 0.5 * nrm * nrm + d - d * v + d * log(v)) {
      g = d * v;
      break;
    }
  }
  double lg = log(g);
  if (alpha < 1.0) lg += log(boost_u) / alpha;
  return lg;
}

__global__ void __launch_bounds__(kMaskThreads)
sample_masks_kernel(const MaskParams p, const unsigned long long seed, unsigned long long* draw_counter,
                    unsigned int* done_counter, int64_t* __restrict__ task_masks,
                    int64_t* __restrict__ ids_keep, int64_t* __restrict__ ids_restore) {
  extern __shared__ unsigned long long s_key[];          // [n_all] sort keys
  uint32_t* s_noise = reinterpret_cast<uint32_t*>(s_key + p.n_all);  // [n_all] per-task noise
  __shared__ double s_lg[kMaskMaxTasks];
  __shared__ int s_per_task[kMaskMaxTasks];
  __shared__ int s_start[kMaskMaxTasks + 1];

  const uint32_t b = blockIdx.x;
  const unsigned long long draw = *draw_counter;
  const uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  const uint32_t d0 = static_cast<uint32_t>(draw), d1 = static_cast<uint32_t>(draw >> 32);

  if (threadIdx.x < p.n_tasks) {
    const int t = threadIdx.x;
    double alpha = p.alphas[t];
    if (p.uniform_tasks) {
      // uniformly chosen non-empty subset of the tasks; row index as in itertools.product([0,1], repeat=T)[1:]
      const Philox4 r = philox4x32_10(k0, k1, d0, d1, b, 1u << 31);
      const uint32_t n_sub = (1u << p.n_tasks) - 1u;
      const uint32_t pick = 1u + static_cast<uint32_t>((static_cast<unsigned long long>(r.x) * n_sub) >> 32);
      const bool on = (pick >> (p.n_tasks - 1 - t)) & 1u;
      alpha = (on ? alpha : 0.0) + 1e-5;
    }
    s_lg[t] = log_gamma_variate(alpha, k0, k1, d0, d1, b, t);
  }
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int t = 0; t < p.n_tasks; ++t) {
      s_start[t] = acc;
      acc += p.counts[t];
    }
    s_start[p.n_tasks] = acc;
  }
  __syncthreads();
  if (threadIdx.x < p.n_tasks) {
    double mx = s_lg[0];
    for (int t = 1; t < p.n_tasks; ++t) mx = fmax(mx, s_lg[t]);
    double sum = 0.0;
    for (int t = 0; t < p.n_tasks; ++t) sum += exp(s_lg[t] - mx);
    const float share = static_cast<float>(exp(s_lg[threadIdx.x] - mx) / sum);
    s_per_task[threadIdx.x] = static_cast<int>(rintf(share * static_cast<float>(p.n_enc)));  // torch.round: half to even
  }
  // noise words
  for (int i = threadIdx.x; i < p.n_all; i += kMaskThreads) {
    const Philox4 r = philox4x32_10(k0, k1, d0, d1, b, static_cast<uint32_t>(i));
    s_noise[i] = r.x;
    s_key[i] = r.y;  // low word of the sort key; the mask bit is OR-ed in below
  }
  __syncthreads();
  // rank inside the task -> task-level mask -> key
  for (int i = threadIdx.x; i < p.n_all; i += kMaskThreads) {
    int t = 0;
    while (i >= s_start[t + 1]) ++t;
    const uint32_t mine = s_noise[i];
    int rank = 0;
    for (int j = s_start[t]; j < s_start[t + 1]; ++j) {
      const uint32_t o = s_noise[j];
      rank += (o < mine || (o == mine && j < i)) ? 1 : 0;
    }
    const unsigned long long masked = rank >= s_per_task[t] ? 1ull : 0ull;
    s_key[i] |= masked << 32;
  }
  __syncthreads();
  // global order
  int64_t* tm = task_masks + static_cast<long long>(b) * p.n_all;
  int64_t* keep = ids_keep + static_cast<long long>(b) * p.n_enc;
  int64_t* restore = ids_restore + static_cast<long long>(b) * p.n_all;
  for (int i = threadIdx.x; i < p.n_all; i += kMaskThreads) {
    const unsigned long long mine = s_key[i];
    int pos = 0;
    for (int j = 0; j < p.n_all; ++j) {
      const unsigned long long o = s_key[j];
      pos += (o < mine || (o == mine && j < i)) ? 1 : 0;
    }
    restore[i] = pos;
    tm[i] = pos >= p.n_enc ? 1 : 0;
    if (pos < p.n_enc) keep[pos] = i;
  }
  // the last CTA to finish advances the draw counter (every CTA has read it by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(done_counter, 1u);
    if (done == gridDim.x - 1) {
      *done_counter = 0u;
      *draw_counter = draw + 1ull;
      __threadfence();
    }
  }
}

}  // namespace mb200

using namespace mb200;

extern "C" int mb_sample_masks(uint64_t seed, uint64_t* draw_counter, uint32_t* done_counter,
                               const int32_t* task_counts, const float* alphas, int32_t n_tasks, int64_t batch,
                               int64_t n_encoded, int32_t uniform_tasks, int64_t* task_masks, int64_t* ids_keep,
                               int64_t* ids_restore, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(draw_counter && done_counter && task_counts && alphas && task_masks && ids_keep && ids_restore,
             "mb_sample_masks: null pointer");
  MB_REQUIRE(n_tasks >= 1 && n_tasks <= kMaskMaxTasks, "mb_sample_masks: %d tasks unsupported (1..%d)", n_tasks,
             kMaskMaxTasks);
  MB_REQUIRE(batch > 0 && batch < (1ll << 31), "mb_sample_masks: batch out of range");
  MaskParams p;
  p.n_tasks = n_tasks;
  p.n_all = 0;
  for (int t = 0; t < kMaskMaxTasks; ++t) {
    p.counts[t] = t < n_tasks ? task_counts[t] : 0;
    p.alphas[t] = t < n_tasks ? alphas[t] : 0.f;
    if (t < n_tasks) {
      MB_REQUIRE(task_counts[t] > 0, "mb_sample_masks: task %d has no tokens", t);
      MB_REQUIRE(alphas[t] > 0.f, "mb_sample_masks: alpha[%d] must be positive", t);
      p.n_all += task_counts[t];
    }
  }
  MB_REQUIRE(n_encoded >= 0 && n_encoded <= p.n_all, "mb_sample_masks: n_encoded %lld outside [0, %d]",
             (long long)n_encoded, p.n_all);
  MB_REQUIRE(p.n_all <= 16384, "mb_sample_masks: %d tokens per sample exceed the shared-memory layout", p.n_all);
  p.n_enc = static_cast<int>(n_encoded);
  p.uniform_tasks = uniform_tasks;
  const size_t smem = static_cast<size_t>(p.n_all) * 12;
  if (smem > 48 * 1024)
    MB_CHECK_CUDA(cudaFuncSetAttribute(sample_masks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
  sample_masks_kernel<<<static_cast<unsigned>(batch), kMaskThreads, smem, stream>>>(
      p, seed, reinterpret_cast<unsigned long long*>(draw_counter), done_counter, task_masks, ids_keep,
      ids_restore);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
