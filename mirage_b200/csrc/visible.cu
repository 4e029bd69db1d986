// mirage_b200/csrc/visible.cu
//
// Visible-token embedding of a MultiMAE pretraining step.  The reference embeds EVERY patch of every modality
// (3 x 256 tokens per sample at cfg 4), concatenates, and only then keeps the ~98 sampled tokens
// (mirage/model.py:352-356 input adapters, :384-388 torch.gather) -- 87 % of the patch-projection work, and of its
// backward, is thrown away.  The masks depend on the token COUNTS only (model.py:168-239), so they are known before
// the adapters run; here the patches of the kept tokens are gathered first and only those rows are projected.
// Rows are laid out exactly as the encoder consumes them, t = b * (n_keep + n_glob) + j:
//
//   visible_rows        ids_keep -> per modality m: row_src[m][t] = source patch index (b * n_tok_m + token) when
//                       row t shows a token of modality m, else -1; row_cls[t] = m, or n_mod + g for global token g
//   embed_rows_init     tok[t] = bias_m + pos_emb_m[token]   |   global_token[g]          (fp32 [T, D])
//   gather_patches32    A_m[t] = the 32 x 32 patch (fp32 for the tf32 projection, bf16 twin for wgrad) or zeros
//   (semseg_patches with a row list: adapters.cu)
//   then per modality ONE GEMM  tok += A_m W_m^T  (zero rows add nothing), and in backward
//   dW_m = dtok^T A_m, and class_colsum: bias / global-token gradients = sums of dtok rows by row class.
//
// Every visible token's value is what the reference computes for it (same K order in the tensor core; the bias and
// position rows are summed before instead of after the accumulator: <= 1 ulp).
#include "../../include/mirage_b200.h"
#include "common.cuh"

namespace mb200 {

constexpr int kVisMaxMods = 4;
constexpr int kVisMaxCls = 8;

struct VisMods {
  int n_mod;
  int start[kVisMaxMods];   // first token of the modality in the concatenated sequence
  int count[kVisMaxMods];   // tokens per sample
};

struct VisInit {
  const float* bias[kVisMaxMods];
  const float* pos[kVisMaxMods];   // [count, D] or null
};

__global__ void __launch_bounds__(256)
visible_rows_kernel(const long long* __restrict__ ids_keep, int B, int n_keep, int n_glob, VisMods vm,
                    int* __restrict__ row_src, int* __restrict__ row_cls) {
  const int n_row = n_keep + n_glob;
  const long long T = static_cast<long long>(B) * n_row;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int b = static_cast<int>(t / n_row), j = static_cast<int>(t % n_row);
  int cls = vm.n_mod + (j - n_keep);
  int src[kVisMaxMods];
#pragma unroll
  for (int m = 0; m < kVisMaxMods; ++m) src[m] = -1;
  if (j < n_keep) {
    const long long id = ids_keep[static_cast<long long>(b) * n_keep + j];
    cls = -1;
#pragma unroll
    for (int m = 0; m < kVisMaxMods; ++m)
      if (m < vm.n_mod && id >= vm.start[m] && id < vm.start[m] + vm.count[m]) {
        cls = m;
        src[m] = b * vm.count[m] + static_cast<int>(id - vm.start[m]);
      }
  }
  row_cls[t] = cls;   // -1: index outside every modality (torch.gather would raise); the row stays zero
#pragma unroll
  for (int m = 0; m < kVisMaxMods; ++m)
    if (m < vm.n_mod) row_src[m * T + t] = src[m];
}

// one warp per row
__global__ void __launch_bounds__(256)
embed_rows_init_kernel(const int* __restrict__ row_src, const int* __restrict__ row_cls, VisMods vm, VisInit vi,
                       const float* __restrict__ glob, float* __restrict__ out, long long T, int D) {
  const long long t = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (t >= T) return;
  const int lane = threadIdx.x & 31;
  const int cls = row_cls[t];
  float* o = out + t * D;
  if (cls < 0) {
    for (int c = lane * 4; c < D; c += 128) *reinterpret_cast<float4*>(o + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  if (cls >= vm.n_mod) {
    const float* g = glob + static_cast<long long>(cls - vm.n_mod) * D;
    for (int c = lane * 4; c < D; c += 128) *reinterpret_cast<float4*>(o + c) = *reinterpret_cast<const float4*>(g + c);
    return;
  }
  const float* bias = vi.bias[cls];
  const float* pos = vi.pos[cls];
  const int tok = row_src[cls * T + t] % vm.count[cls];
  for (int c = lane * 4; c < D; c += 128) {
    float4 v = bias != nullptr ? *reinterpret_cast<const float4*>(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (pos != nullptr) {
      const float4 q = *reinterpret_cast<const float4*>(pos + static_cast<long long>(tok) * D + c);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    *reinterpret_cast<float4*>(o + c) = v;
  }
}

// one warp per row, lane = patch row: 128 contiguous bytes in, 128 (fp32) + 64 (bf16) contiguous bytes out
__global__ void __launch_bounds__(256)
gather_patches32_kernel(const float* __restrict__ img, const int* __restrict__ row_src, float* __restrict__ a_f32,
                        __nv_bfloat16* __restrict__ a_bf16, long long T, int H, int W) {
  const long long t = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (t >= T) return;
  const int lane = threadIdx.x & 31;
  const int src = row_src[t];
  float4 v[8];
  if (src < 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    const int gw = W / 32, gh = H / 32;
    const int b = src / (gh * gw), tk = src % (gh * gw);
    const int nh = tk / gw, nw = tk % gw;
    const float* p = img + (static_cast<long long>(b) * H + nh * 32 + lane) * W + nw * 32;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(p + i * 4);
  }
  if (a_f32 != nullptr) {
    float* o = a_f32 + t * 1024 + lane * 32;
#pragma unroll
    for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(o + i * 4) = v[i];
  }
  if (a_bf16 != nullptr) {
    __nv_bfloat16* o = a_bf16 + t * 1024 + lane * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 pk;
      pk.x = pack_bf16x2(v[2 * i].x, v[2 * i].y);
      pk.y = pack_bf16x2(v[2 * i].z, v[2 * i].w);
      pk.z = pack_bf16x2(v[2 * i + 1].x, v[2 * i + 1].y);
      pk.w = pack_bf16x2(v[2 * i + 1].z, v[2 * i + 1].w);
      *reinterpret_cast<uint4*>(o + i * 8) = pk;
    }
  }
}

// part[blk][cls][D] = sum over the block's rows of class cls of dy[row, :]   (thread = 4 columns; the class is
// uniform over the block, so the shared accumulators have one owner per element: no atomics, fixed order)
__global__ void __launch_bounds__(256)
class_colsum_partial_kernel(const float* __restrict__ dy, const int* __restrict__ row_cls,
                            float* __restrict__ part, long long T, int D, int n_cls, int rows_per_block) {
  extern __shared__ float s_acc[];   // [n_cls][D]
  for (int i = threadIdx.x; i < n_cls * D; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = (r0 + rows_per_block < T) ? r0 + rows_per_block : T;
  const int c = threadIdx.x * 4;
  if (c < D) {
    for (long long r = r0; r < r1; r += 4) {
      float4 v[4];
      int cls[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool in = r + u < r1;
        cls[u] = in ? row_cls[r + u] : -1;
        v[u] = in ? *reinterpret_cast<const float4*>(dy + (r + u) * D + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (cls[u] < 0 || cls[u] >= n_cls) continue;
        float* a = s_acc + cls[u] * D + c;
        a[0] += v[u].x; a[1] += v[u].y; a[2] += v[u].z; a[3] += v[u].w;
      }
    }
  }
  __syncthreads();
  float* o = part + static_cast<long long>(blockIdx.x) * n_cls * D;
  for (int i = threadIdx.x; i < n_cls * D; i += blockDim.x) o[i] = s_acc[i];
}

__global__ void __launch_bounds__(256)
class_colsum_final_kernel(const float* __restrict__ part, float* __restrict__ out, int n_blocks, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int b = 0; b < n_blocks; ++b) acc += part[static_cast<long long>(b) * n + i];
  out[i] = acc;
}

static int colsum_blocks(long long T) {
  long long nb = (T + 63) / 64;
  const long long cap = 296;   // fixed (not sm_count(): the workspace query and the launch must agree)
  return static_cast<int>(nb < 1 ? 1 : (nb > cap ? cap : nb));
}

}  // namespace mb200

using namespace mb200;

extern "C" {

int mb_visible_rows(const int64_t* ids_keep, int64_t batch, int64_t n_keep, int64_t n_global, int32_t n_modalities,
                    const int32_t* starts, const int32_t* counts, int32_t* row_src, int32_t* row_cls, void* stream) {
  MB_REQUIRE(ids_keep && starts && counts && row_src && row_cls, "mb_visible_rows: null pointer");
  MB_REQUIRE(n_modalities >= 1 && n_modalities <= kVisMaxMods && n_global >= 0 &&
                 n_modalities + n_global <= kVisMaxCls,
             "mb_visible_rows: %d modalities + %lld global tokens unsupported (<= %d, <= %d row classes)",
             n_modalities, (long long)n_global, kVisMaxMods, kVisMaxCls);
  VisMods vm;
  vm.n_mod = n_modalities;
  for (int m = 0; m < kVisMaxMods; ++m) {
    vm.start[m] = m < n_modalities ? starts[m] : 0;
    vm.count[m] = m < n_modalities ? counts[m] : 1;
    if (m < n_modalities)
      MB_REQUIRE(counts[m] > 0 && batch * counts[m] < (1ll << 31), "mb_visible_rows: modality %d out of range", m);
  }
  const long long T = batch * (n_keep + n_global);
  if (T == 0) return 0;
  visible_rows_kernel<<<static_cast<unsigned>((T + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ids_keep), static_cast<int>(batch), static_cast<int>(n_keep),
      static_cast<int>(n_global), vm, row_src, row_cls);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mb_embed_rows_init(const int32_t* row_src, const int32_t* row_cls, int32_t n_modalities, const int32_t* counts,
                       const float* const* bias, const float* const* pos, const float* global_tokens, float* out,
                       int64_t rows, int64_t dim, void* stream) {
  MB_REQUIRE(row_src && row_cls && counts && bias && pos && out, "mb_embed_rows_init: null pointer");
  MB_REQUIRE(n_modalities >= 1 && n_modalities <= kVisMaxMods, "mb_embed_rows_init: %d modalities unsupported",
             n_modalities);
  MB_REQUIRE(dim % 4 == 0, "mb_embed_rows_init: dim must be a multiple of 4");
  VisMods vm;
  VisInit vi;
  vm.n_mod = n_modalities;
  for (int m = 0; m < kVisMaxMods; ++m) {
    vm.start[m] = 0;
    vm.count[m] = m < n_modalities ? counts[m] : 1;
    vi.bias[m] = m < n_modalities ? bias[m] : nullptr;
    vi.pos[m] = m < n_modalities ? pos[m] : nullptr;
  }
  if (rows == 0) return 0;
  embed_rows_init_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      row_src, row_cls, vm, vi, global_tokens, out, rows, static_cast<int>(dim));
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mb_gather_patches32(const float* images, const int32_t* row_src, float* a_f32, void* a_bf16, int64_t rows,
                        int64_t height, int64_t width, void* stream) {
  MB_REQUIRE(images && row_src && (a_f32 || a_bf16), "mb_gather_patches32: null pointer");
  MB_REQUIRE(height % 32 == 0 && width % 32 == 0, "mb_gather_patches32: image sides must be multiples of 32");
  MB_REQUIRE((reinterpret_cast<uintptr_t>(images) & 15) == 0, "mb_gather_patches32: images must be 16-byte aligned");
  if (rows == 0) return 0;
  gather_patches32_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      images, row_src, a_f32, reinterpret_cast<__nv_bfloat16*>(a_bf16), rows, static_cast<int>(height),
      static_cast<int>(width));
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int64_t mb_class_colsum_workspace(int64_t rows, int64_t dim, int32_t n_classes) {
  return static_cast<int64_t>(colsum_blocks(rows)) * n_classes * dim * static_cast<int64_t>(sizeof(float));
}

int mb_class_colsum(const float* dy, const int32_t* row_cls, float* out, void* workspace, int64_t rows, int64_t dim,
                    int32_t n_classes, void* stream) {
  MB_REQUIRE(dy && row_cls && out && workspace, "mb_class_colsum: null pointer");
  MB_REQUIRE(dim % 4 == 0 && dim <= 1024 && n_classes >= 1 && n_classes <= kVisMaxCls,
             "mb_class_colsum: dim %lld / %d classes unsupported", (long long)dim, n_classes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nb = colsum_blocks(rows);
  const int rpb = static_cast<int>((rows + nb - 1) / nb);
  const size_t smem = static_cast<size_t>(n_classes) * dim * sizeof(float);
  if (smem > 48 * 1024)
    MB_CHECK_CUDA(cudaFuncSetAttribute(class_colsum_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
  float* part = reinterpret_cast<float*>(workspace);
  class_colsum_partial_kernel<<<nb, 256, smem, st>>>(dy, row_cls, part, rows, static_cast<int>(dim), n_classes, rpb);
  MB_CHECK_CUDA(cudaGetLastError());
  const int n = n_classes * static_cast<int>(dim);
  class_colsum_final_kernel<<<(n + 255) / 256, 256, 0, st>>>(part, out, nb, n);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
