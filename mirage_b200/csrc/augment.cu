// mirage_b200/csrc/augment.cu
//
// Device-side input pipeline of MultiMAE pretraining: the reference decodes uint8
// .npy files and augments every sample on CPU worker processes (mutils/datasets_pretrain.py:18-83,
// :172-185), then ships fp32 images over PCIe.  Here the RAW uint8 batch travels (4x fewer bytes) and one
// gather kernel per modality applies, in the reference's order:
//     uint8 -> [0,1] float (:178)  ->  horizontal flip (:43-44)  ->  intensity shift + clip (:45-50, images only)
//     ->  affine warp, bilinear, fill 0 (:54-67; torchvision F.affine on tensors = affine_grid +
//         grid_sample(align_corners=False, zeros padding) times the equally warped all-ones mask)  ->  nearest resize of the layer map to its input size (:68-78)
// The per-sample parameters (flip, shift, inverse affine matrix in torchvision's centred convention) are
// drawn on the host and passed as a small table, so the random stream stays the caller's business.
//
// params[b] = {flip, shift, m00, m01, m02, m10, m11, m12}: source position of output pixel (x, y) is
//     px = m00 (x - cx) + m01 (y - cy) + m02 + cx,   py = m10 (x - cx) + m11 (y - cy) + m12 + cy,
//     cx = (W - 1) / 2, cy = (H - 1) / 2.
// HBM-bound: 1 byte read (gathered, L1/L2-resident neighbourhoods) + 4 (image) or 8/16 (label) bytes written
// per output pixel.
#include "../../include/mirage_b200.h"
#include "common.cuh"

namespace mb200 {

constexpr int kAugThreads = 256;

struct AugRow {
  float flip, shift, m00, m01, m02, m10, m11, m12;
};

template <bool LABELS>
__device__ __forceinline__ float aug_fetch(const uint8_t* __restrict__ img, int H, int W, int y, int x, bool flip,
                                           float shift, float& inside) {
  if (x < 0 || x >= W || y < 0 || y >= H) {  // grid_sample padding_mode = zeros
    inside = 0.f;
    return 0.f;
  }
  inside = 1.f;
  const uint8_t raw = img[y * W + (flip ? W - 1 - x : x)];
  if (LABELS) return static_cast<float>(raw);
  return fminf(fmaxf(static_cast<float>(raw) / 255.0f + shift, 0.f), 1.f);
}

template <bool LABELS>
__device__ __forceinline__ float aug_sample(const uint8_t* __restrict__ img, int H, int W, float ox, float oy,
                                            const AugRow& r, bool identity) {
  const bool flip = r.flip != 0.f;
  float i00, i01, i10, i11;
  if (identity) return aug_fetch<LABELS>(img, H, W, static_cast<int>(oy), static_cast<int>(ox), flip, r.shift, i00);
  const float cx = 0.5f * (W - 1), cy = 0.5f * (H - 1);
  const float dx = ox - cx, dy = oy - cy;
  const float px = r.m00 * dx + r.m01 * dy + r.m02 + cx;
  const float py = r.m10 * dx + r.m11 * dy + r.m12 + cy;
  const float fx = floorf(px), fy = floorf(py);
  const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
  const float wx = px - fx, wy = py - fy;
  const float v00 = aug_fetch<LABELS>(img, H, W, y0, x0, flip, r.shift, i00);
  const float v01 = aug_fetch<LABELS>(img, H, W, y0, x0 + 1, flip, r.shift, i01);
  const float v10 = aug_fetch<LABELS>(img, H, W, y0 + 1, x0, flip, r.shift, i10);
  const float v11 = aug_fetch<LABELS>(img, H, W, y0 + 1, x0 + 1, flip, r.shift, i11);
  // grid_sample's bilinear form: sum of the four corner values times their area weights
  const float w00 = (1.f - wx) * (1.f - wy), w01 = wx * (1.f - wy), w10 = (1.f - wx) * wy, w11 = wx * wy;
  const float val = v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11;
  // torchvision's fill handling (F_t._apply_grid_transform): a channel of ones is warped along with the image and
  // the result is img * mask + (1 - mask) * fill -- with fill = 0 the border is attenuated TWICE
  const float mask = i00 * w00 + i01 * w01 + i10 * w10 + i11 * w11;
  return val * mask;
}

__device__ __forceinline__ bool aug_is_identity(const AugRow& r) {
  return r.m00 == 1.f && r.m01 == 0.f && r.m02 == 0.f && r.m10 == 0.f && r.m11 == 1.f && r.m12 == 0.f;
}

__global__ void __launch_bounds__(kAugThreads)
augment_image_kernel(const uint8_t* __restrict__ src, const AugRow* __restrict__ params, float* __restrict__ out,
                     int H, int W) {
  const int b = blockIdx.y;
  const AugRow r = params[b];
  const bool identity = aug_is_identity(r);
  const uint8_t* img = src + static_cast<long long>(b) * H * W;
  float* o = out + static_cast<long long>(b) * H * W;
  const int idx = (blockIdx.x * kAugThreads + threadIdx.x) * 4;  // four consecutive pixels of a row (W % 4 == 0)
  if (idx >= H * W) return;
  const int y = idx / W, x = idx % W;
  float4 v;
  v.x = aug_sample<false>(img, H, W, static_cast<float>(x), static_cast<float>(y), r, identity);
  v.y = aug_sample<false>(img, H, W, static_cast<float>(x + 1), static_cast<float>(y), r, identity);
  v.z = aug_sample<false>(img, H, W, static_cast<float>(x + 2), static_cast<float>(y), r, identity);
  v.w = aug_sample<false>(img, H, W, static_cast<float>(x + 3), static_cast<float>(y), r, identity);
  *reinterpret_cast<float4*>(o + idx) = v;
}

__global__ void __launch_bounds__(kAugThreads)
augment_labels_kernel(const uint8_t* __restrict__ src, const AugRow* __restrict__ params, int64_t* __restrict__ out,
                      int H, int W, int OH, int OW) {
  const int b = blockIdx.y;
  const AugRow r = params[b];
  const bool identity = aug_is_identity(r);
  const uint8_t* img = src + static_cast<long long>(b) * H * W;
  const int idx = blockIdx.x * kAugThreads + threadIdx.x;
  if (idx >= OH * OW) return;
  const int y = idx / OW, x = idx % OW;
  // nearest resize (torch 'nearest'): source index = floor(dst * in / out)
  const int sy = min(H - 1, static_cast<int>(floorf(y * (static_cast<float>(H) / OH))));
  const int sx = min(W - 1, static_cast<int>(floorf(x * (static_cast<float>(W) / OW))));
  const float v = aug_sample<true>(img, H, W, static_cast<float>(sx), static_cast<float>(sy), r, identity);
  // torchvision interpolates integer images in float and rounds back (half to even)
  out[static_cast<long long>(b) * OH * OW + idx] = static_cast<int64_t>(rintf(v));
}

}  // namespace mb200

using namespace mb200;

extern "C" {

int mb_augment_image(const uint8_t* src, const float* params, float* out, int64_t batch, int32_t height,
                     int32_t width, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(src && params && out, "mb_augment_image: null pointer");
  MB_REQUIRE(batch > 0 && batch <= 65535 && height > 0 && width > 0 && width % 4 == 0,
             "mb_augment_image: unsupported geometry (batch %lld, %d x %d)", (long long)batch, height, width);
  const int quads = height * width / 4;
  dim3 grid(static_cast<unsigned>((quads + kAugThreads - 1) / kAugThreads), static_cast<unsigned>(batch));
  augment_image_kernel<<<grid, kAugThreads, 0, stream>>>(src, reinterpret_cast<const AugRow*>(params), out, height,
                                                         width);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mb_augment_labels(const uint8_t* src, const float* params, int64_t* out, int64_t batch, int32_t height,
                      int32_t width, int32_t out_height, int32_t out_width, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(src && params && out, "mb_augment_labels: null pointer");
  MB_REQUIRE(batch > 0 && batch <= 65535 && height > 0 && width > 0 && out_height > 0 && out_width > 0,
             "mb_augment_labels: unsupported geometry");
  const int px = out_height * out_width;
  dim3 grid(static_cast<unsigned>((px + kAugThreads - 1) / kAugThreads), static_cast<unsigned>(batch));
  augment_labels_kernel<<<grid, kAugThreads, 0, stream>>>(src, reinterpret_cast<const AugRow*>(params), out, height,
                                                          width, out_height, out_width);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
