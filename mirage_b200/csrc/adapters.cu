// mirage_b200/csrc/adapters.cu
//
// Indexing kernels around the input / output adapters (all HBM- or L2-bound, 16-byte accesses):
//   semseg_patches / class_emb_grad   SemSegInputAdapter embedding lookup fused with patch extraction
//                                     (mirage/input_adapters.py:226-229) and its backward
//   dec_assemble fwd / bwd            SpatialOutputAdapter.get_queries_and_context
//                                     (mirage/output_adapters.py:188-246)
//   patchify_cast                     image-layout fp32 gradient -> token-layout bf16 GEMM operand
//                                     (backward of the un-patchify at output_adapters.py:291-294)
#include "../../include/mirage_b200.h"
#include "common.cuh"

namespace mb200 {

__device__ __forceinline__ float4 ldf4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void stf4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ------------------------------------------------------------------------------------------
// A[m, c*P*Q + ph*Q + pw] = class_emb[labels[b, nh*P + ph, nw*Q + pw], c]      (bf16)
// one block per patch m = (b, nh, nw); the 13x64 table and the patch's labels sit in shared memory;
// thread -> 8 consecutive pw (one 16-byte store), consecutive threads -> consecutive addresses.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
semseg_patches_kernel(const long long* __restrict__ labels, const __nv_bfloat16* __restrict__ table,
                      __nv_bfloat16* __restrict__ out, int H, int W, int P, int Q, int n_cls, int E,
                      const int* __restrict__ row_src) {
  extern __shared__ uint8_t sm[];
  __nv_bfloat16* s_tab = reinterpret_cast<__nv_bfloat16*>(sm);             // [n_cls * E]
  int* s_lab = reinterpret_cast<int*>(sm + ((n_cls * E * 2 + 15) & ~15));  // [P * Q]
  const int gw = W / Q, gh = H / P;
  // row list (visible-token embedding, visible.cu): output row blockIdx.x holds source patch row_src[blockIdx.x],
  // or zeros when that is negative (a row that belongs to another modality)
  const int m = row_src != nullptr ? row_src[blockIdx.x] : static_cast<int>(blockIdx.x);
  if (m < 0) {
    uint4* z = reinterpret_cast<uint4*>(out + (long long)blockIdx.x * E * P * Q);
    for (int v = threadIdx.x; v < E * P * Q / 8; v += blockDim.x) z[v] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const int b = m / (gh * gw), t = m % (gh * gw);
  const int nh = t / gw, nw = t % gw;
  for (int i = threadIdx.x; i < n_cls * E; i += blockDim.x) s_tab[i] = table[i];
  for (int i = threadIdx.x; i < P * Q; i += blockDim.x) {
    const int ph = i / Q, pw = i % Q;
    long long l = labels[((long long)b * H + nh * P + ph) * W + nw * Q + pw];
    s_lab[i] = (int)(l < 0 ? 0 : (l >= n_cls ? n_cls - 1 : l));
  }
  __syncthreads();
  const int q8 = Q / 8;
  const int nvec = E * P * q8;
  __nv_bfloat16* orow = out + (long long)blockIdx.x * E * P * Q;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const int c = v / (P * q8), r = v % (P * q8);
    const int ph = r / q8, pw0 = (r % q8) * 8;
    __align__(16) __nv_bfloat16 vals[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) vals[i] = s_tab[s_lab[ph * Q + pw0 + i] * E + c];
    *reinterpret_cast<uint4*>(orow + (long long)c * P * Q + ph * Q + pw0) =
        *reinterpret_cast<const uint4*>(vals);
  }
}

// dE[cls, c] += sum over (m, ph, pw) with label == cls of dA[m, c*P*Q + ph*Q + pw]
// A block walks `patches_per_block` patches.  Per patch the [E, P*Q] bf16 gradient block is staged in
// shared memory (coalesced 16-byte loads, rows padded against bank conflicts); then warp w owns the
// channels (w & 1) * 32 + lane and the pixel quarter w >> 1: all lanes of a warp look at the SAME pixel,
// so the label is warp-uniform and lane c adds into acc[quarter][label][c] -- its own column: no
// atomics, no conflicts (the first version issued 8 shared-memory atomics per thread per vector, 847 us
// for the cfg-4 batch against ~90 us of HBM time).  dE must be zero-filled by the caller.
__global__ void __launch_bounds__(256)
class_emb_grad_kernel(const long long* __restrict__ labels, const __nv_bfloat16* __restrict__ dA,
                      float* __restrict__ dE, int n_patches, int H, int W, int P, int Q, int n_cls,
                      int E, int patches_per_block, const int* __restrict__ row_src) {
  extern __shared__ uint8_t sm[];
  const int PQ = P * Q;
  const int row_words = PQ / 2 + 1;                                   // bf16 pairs per channel row, +1 pad
  float* s_acc = reinterpret_cast<float*>(sm);                        // [4][n_cls][E]
  uint32_t* s_val = reinterpret_cast<uint32_t*>(s_acc + 4 * n_cls * E);  // [E][row_words]
  int* s_lab = reinterpret_cast<int*>(s_val + E * row_words);         // [PQ]
  for (int i = threadIdx.x; i < 4 * n_cls * E; i += blockDim.x) s_acc[i] = 0.f;
  const int gw = W / Q, gh = H / P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp >> 1;
  const int px0 = quarter * (PQ / 4), px1 = px0 + PQ / 4;
  const int m0 = blockIdx.x * patches_per_block;
  for (int mo = m0; mo < min(m0 + patches_per_block, n_patches); ++mo) {
    // row list: gradient row mo belongs to source patch row_src[mo] (negative: not of this modality, skip)
    const int m = row_src != nullptr ? row_src[mo] : mo;
    if (m < 0) continue;   // block-uniform
    const int b = m / (gh * gw), t = m % (gh * gw);
    const int nh = t / gw, nw = t % gw;
    __syncthreads();  // previous patch fully consumed (also orders the zero-fill)
    for (int i = threadIdx.x; i < PQ; i += blockDim.x) {
      const int ph = i / Q, pw = i % Q;
      long long l = labels[((long long)b * H + nh * P + ph) * W + nw * Q + pw];
      s_lab[i] = (int)(l < 0 ? 0 : (l >= n_cls ? n_cls - 1 : l));
    }
    const uint4* arow = reinterpret_cast<const uint4*>(dA + (long long)mo * E * PQ);
    const int vec_per_row = PQ / 8;
    for (int v = threadIdx.x; v < E * vec_per_row; v += blockDim.x) {
      const int c = v / vec_per_row, j = v % vec_per_row;
      const uint4 pk = arow[v];
      uint32_t* d = s_val + c * row_words + j * 4;
      d[0] = pk.x; d[1] = pk.y; d[2] = pk.z; d[3] = pk.w;
    }
    __syncthreads();
    for (int c = (warp & 1) * 32 + lane; c < E; c += 64) {
      const uint32_t* vrow = s_val + c * row_words;
      for (int px = px0; px < px1; px += 2) {
        const float2 f = unpack_bf16x2(vrow[px >> 1]);
        float* a0 = s_acc + (quarter * n_cls + s_lab[px]) * E + c;
        float* a1 = s_acc + (quarter * n_cls + s_lab[px + 1]) * E + c;
        *a0 += f.x;
        *a1 += f.y;   // (a1 may alias a0: plain sequential adds by the same thread)
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_cls * E; i += blockDim.x) {
    const float v = s_acc[i] + s_acc[n_cls * E + i] + s_acc[2 * n_cls * E + i] + s_acc[3 * n_cls * E + i];
    if (v != 0.f) atomicAdd(&dE[i], v);
  }
}

// ------------------------------------------------------------------------------------------
// decoder queries / context assembly; one warp per output row, float4 lanes
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dec_assemble_fwd_kernel(const float* __restrict__ ctx, const float* __restrict__ mask_token,
                        const float* __restrict__ emb, const long long* __restrict__ ids_keep,
                        const long long* __restrict__ ids_restore, float* __restrict__ q,
                        float* __restrict__ c, int B, int n_vis, int n_glob, int n_all, int q_start,
                        int n_q, int Dd) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int n_ctx = n_vis + n_glob;
  const long long total = (long long)B * (n_q + n_ctx);
  if (row >= total) return;
  const int b = (int)(row / (n_q + n_ctx));
  const int i = (int)(row % (n_q + n_ctx));
  const float* src;
  const float* e = nullptr;
  float* dst;
  if (i < n_q) {
    const int pos = q_start + i;
    const long long r = ids_restore[(long long)b * n_all + pos];
    src = (r >= 0 && r < n_vis) ? ctx + ((long long)b * n_ctx + r) * Dd : mask_token;
    e = emb + (long long)pos * Dd;
    dst = q + ((long long)b * n_q + i) * Dd;
  } else {
    const int j = i - n_q;
    dst = c + ((long long)b * n_ctx + j) * Dd;
    if (j < n_vis) {
      long long pos = ids_keep[(long long)b * n_vis + j];
      pos = pos < 0 ? 0 : (pos >= n_all ? n_all - 1 : pos);
      const long long r = ids_restore[(long long)b * n_all + pos];
      src = (r >= 0 && r < n_vis) ? ctx + ((long long)b * n_ctx + r) * Dd : mask_token;
      e = emb + pos * Dd;
    } else {
      src = ctx + ((long long)b * n_ctx + j) * Dd;  // global tokens: no task / positional embedding
    }
  }
  for (int col = lane * 4; col < Dd; col += 128) {
    float4 v = ldf4(src + col);
    if (e) {
      const float4 w = ldf4(e + col);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    stf4(dst + col, v);
  }
}

// d_ctx[b, r] = dc[b, r] + (token position of visible row r lies in this task ? dq[b, pos - q_start] : 0)
// (ids_restore is the inverse permutation of the shuffle that produced ids_keep)
__global__ void __launch_bounds__(256)
dec_assemble_bwd_ctx_kernel(const float* __restrict__ dq, const float* __restrict__ dc,
                            const long long* __restrict__ ids_keep, float* __restrict__ dctx, int B,
                            int n_vis, int n_glob, int n_all, int q_start, int n_q, int Dd) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int n_ctx = n_vis + n_glob;
  if (row >= (long long)B * n_ctx) return;
  const int b = (int)(row / n_ctx), j = (int)(row % n_ctx);
  const float* a = dc + row * Dd;
  const float* qsrc = nullptr;
  if (j < n_vis) {
    const long long pos = ids_keep[(long long)b * n_vis + j];
    if (pos >= q_start && pos < q_start + n_q) qsrc = dq + ((long long)b * n_q + (pos - q_start)) * Dd;
  }
  for (int col = lane * 4; col < Dd; col += 128) {
    float4 v = ldf4(a + col);
    if (qsrc) {
      const float4 w = ldf4(qsrc + col);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    stf4(dctx + row * Dd + col, v);
  }
}

// one block per token position p: demb[p] = sum_b (dq if p in task range) + (visible ? dc[b, r] : 0);
// dmask_part[p - q_start] = sum_b (masked ? dq[b, p - q_start] : 0)
__global__ void dec_assemble_bwd_emb_kernel(const float* __restrict__ dq, const float* __restrict__ dc,
                                            const long long* __restrict__ ids_restore,
                                            float* __restrict__ demb, float* __restrict__ dmask_part,
                                            int B, int n_vis, int n_glob, int n_all, int q_start,
                                            int n_q, int Dd) {
  const int p = blockIdx.x;
  const int n_ctx = n_vis + n_glob;
  const bool in_task = (p >= q_start && p < q_start + n_q);
  // The batch loop is a chain of dependent loads (index -> gradient row) and was latency-bound (194 us for 93 MB
  // at cfg 4, one column per thread, 256 samples in sequence).  Threads now own a float4 of columns, and the
  // 256 / (Dd / 4) thread groups of the block take every G-th sample; the group partials are combined in group
  // order through shared memory (fixed summation order).
  __shared__ float4 s_acc[256], s_accm[256];
  const int nvec = Dd / 4;
  if (nvec <= 256 && 256 % nvec == 0 && blockDim.x == 256) {
    const int G = 256 / nvec;
    const int c4 = threadIdx.x % nvec, g = threadIdx.x / nvec;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), accm = acc;
#pragma unroll 4
    for (int b = g; b < B; b += G) {
      const long long r = ids_restore[(long long)b * n_all + p];
      const bool vis = (r >= 0 && r < n_vis);
      if (in_task) {
        const float4 q = ldf4(dq + ((long long)b * n_q + (p - q_start)) * Dd + c4 * 4);
        acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
        if (!vis) { accm.x += q.x; accm.y += q.y; accm.z += q.z; accm.w += q.w; }
      }
      if (vis) {
        const float4 c = ldf4(dc + ((long long)b * n_ctx + r) * Dd + c4 * 4);
        acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w;
      }
    }
    s_acc[threadIdx.x] = acc;
    s_accm[threadIdx.x] = accm;
    __syncthreads();
    if (g == 0) {
      for (int k = 1; k < G; ++k) {
        const float4 a = s_acc[k * nvec + c4], m = s_accm[k * nvec + c4];
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
        accm.x += m.x; accm.y += m.y; accm.z += m.z; accm.w += m.w;
      }
      stf4(demb + (long long)p * Dd + c4 * 4, acc);
      if (in_task) stf4(dmask_part + (long long)(p - q_start) * Dd + c4 * 4, accm);
    }
    return;
  }
  for (int col = threadIdx.x; col < Dd; col += blockDim.x) {
    float acc = 0.f, accm = 0.f;
    for (int b = 0; b < B; ++b) {
      const long long r = ids_restore[(long long)b * n_all + p];
      const bool vis = (r >= 0 && r < n_vis);
      if (in_task) {
        const float g = dq[((long long)b * n_q + (p - q_start)) * Dd + col];
        acc += g;
        if (!vis) accm += g;
      }
      if (vis) acc += dc[((long long)b * n_ctx + r) * Dd + col];
    }
    demb[(long long)p * Dd + col] = acc;
    if (in_task) dmask_part[(long long)(p - q_start) * Dd + col] = accm;
  }
}

__global__ void rowsum_small_kernel(const float* __restrict__ part, float* __restrict__ out, int rows,
                                    int Dd) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= Dd) return;
  float acc = 0.f;
  for (int r = 0; r < rows; ++r) acc += part[(long long)r * Dd + col];
  out[col] = acc;
}

// ------------------------------------------------------------------------------------------
// image [B, C, gh*ph, gw*pw] fp32 -> tokens [B*gh*gw, C*ph*pw] bf16
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
patchify_cast_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, long long n_vec,
                     int C, int ph, int pw, int gh, int gw) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vec) return;
  const int pw8 = pw / 8;
  const int per_tok = C * ph * pw8;
  const long long m = v / per_tok;
  const int r = (int)(v % per_tok);
  const int c = r / (ph * pw8), r2 = r % (ph * pw8);
  const int py = r2 / pw8, px = (r2 % pw8) * 8;
  const int per_img = gh * gw;
  const long long b = m / per_img;
  const int t = (int)(m % per_img);
  const int nh = t / gw, nw = t % gw;
  const long long W = (long long)gw * pw, H = (long long)gh * ph;
  const float* s = img + ((b * C + c) * H + (long long)nh * ph + py) * W + (long long)nw * pw + px;
  const float4 a = ldf4(s), d = ldf4(s + 4);
  uint4 pk;
  pk.x = pack_bf16x2(a.x, a.y);
  pk.y = pack_bf16x2(a.z, a.w);
  pk.z = pack_bf16x2(d.x, d.y);
  pk.w = pack_bf16x2(d.z, d.w);
  *reinterpret_cast<uint4*>(out + m * ((long long)C * ph * pw) + (long long)c * ph * pw + py * pw + px) = pk;
}

}  // namespace mb200

using namespace mb200;

extern "C" {

int mb_semseg_patches(const int64_t* labels, const void* class_emb_bf16, void* out, int64_t batch,
                      int64_t height, int64_t width, int32_t patch_h, int32_t patch_w,
                      int32_t n_classes, int32_t emb_dim, void* stream) {
  return mb_semseg_patches_rows(labels, class_emb_bf16, out, nullptr, 0, batch, height, width, patch_h, patch_w,
                                n_classes, emb_dim, stream);
}

int mb_semseg_patches_rows(const int64_t* labels, const void* class_emb_bf16, void* out, const int32_t* row_src,
                           int64_t n_rows, int64_t batch, int64_t height, int64_t width, int32_t patch_h,
                           int32_t patch_w, int32_t n_classes, int32_t emb_dim, void* stream) {
  MB_REQUIRE(labels && class_emb_bf16 && out, "mb_semseg_patches: null pointer");
  MB_REQUIRE(patch_w % 8 == 0, "mb_semseg_patches: patch width must be a multiple of 8");
  MB_REQUIRE(height % patch_h == 0 && width % patch_w == 0, "mb_semseg_patches: ragged patches");
  const long long m = row_src != nullptr ? n_rows : batch * (height / patch_h) * (width / patch_w);
  if (m == 0) return 0;
  const size_t smem = ((size_t)n_classes * emb_dim * 2 + 15 & ~(size_t)15) + (size_t)patch_h * patch_w * 4;
  semseg_patches_kernel<<<(unsigned)m, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(labels),
      reinterpret_cast<const __nv_bfloat16*>(class_emb_bf16), reinterpret_cast<__nv_bfloat16*>(out),
      (int)height, (int)width, patch_h, patch_w, n_classes, emb_dim, row_src);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mb_class_emb_grad(const int64_t* labels, const void* d_patches_bf16, float* d_class_emb,
                      int64_t batch, int64_t height, int64_t width, int32_t patch_h, int32_t patch_w,
                      int32_t n_classes, int32_t emb_dim, void* stream) {
  return mb_class_emb_grad_rows(labels, d_patches_bf16, d_class_emb, nullptr, 0, batch, height, width, patch_h,
                                patch_w, n_classes, emb_dim, stream);
}

int mb_class_emb_grad_rows(const int64_t* labels, const void* d_patches_bf16, float* d_class_emb,
                           const int32_t* row_src, int64_t n_rows, int64_t batch, int64_t height, int64_t width,
                           int32_t patch_h, int32_t patch_w, int32_t n_classes, int32_t emb_dim, void* stream) {
  MB_REQUIRE(labels && d_patches_bf16 && d_class_emb, "mb_class_emb_grad: null pointer");
  MB_REQUIRE(patch_w % 8 == 0, "mb_class_emb_grad: patch width must be a multiple of 8");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MB_CHECK_CUDA(cudaMemsetAsync(d_class_emb, 0, (size_t)n_classes * emb_dim * sizeof(float), st));
  const long long m = row_src != nullptr ? n_rows : batch * (height / patch_h) * (width / patch_w);
  if (m == 0) return 0;
  MB_REQUIRE((patch_h * patch_w) % 8 == 0, "mb_class_emb_grad: patch area must be a multiple of 8");
  const int ppb = 16;
  const int pq = patch_h * patch_w;
  const size_t smem = (size_t)4 * n_classes * emb_dim * 4 + (size_t)emb_dim * (pq / 2 + 1) * 4 + (size_t)pq * 4;
  MB_REQUIRE(smem <= 200 * 1024, "mb_class_emb_grad: patch / embedding too large for shared memory (%zu B)", smem);
  if (smem > 48 * 1024)
    MB_CHECK_CUDA(cudaFuncSetAttribute(class_emb_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  class_emb_grad_kernel<<<(unsigned)((m + ppb - 1) / ppb), 256, smem, st>>>(
      reinterpret_cast<const long long*>(labels),
      reinterpret_cast<const __nv_bfloat16*>(d_patches_bf16), d_class_emb, (int)m, (int)height,
      (int)width, patch_h, patch_w, n_classes, emb_dim, ppb, row_src);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mb_dec_assemble_fwd(const float* ctx, const float* mask_token, const float* emb,
                        const int64_t* ids_keep, const int64_t* ids_restore, float* queries,
                        float* context, int64_t batch, int64_t n_vis, int64_t n_global, int64_t n_all,
                        int64_t q_start, int64_t n_q, int64_t dim, void* stream) {
  MB_REQUIRE(ctx && mask_token && emb && ids_restore && queries && context && (n_vis == 0 || ids_keep),
             "mb_dec_assemble_fwd: null pointer");
  MB_REQUIRE(dim % 4 == 0, "mb_dec_assemble_fwd: dim must be a multiple of 4");
  MB_REQUIRE(q_start >= 0 && q_start + n_q <= n_all, "mb_dec_assemble_fwd: query range out of bounds");
  const long long rows = batch * (n_q + n_vis + n_global);
  if (rows == 0) return 0;
  dec_assemble_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      ctx, mask_token, emb, reinterpret_cast<const long long*>(ids_keep),
      reinterpret_cast<const long long*>(ids_restore), queries, context, (int)batch, (int)n_vis,
      (int)n_global, (int)n_all, (int)q_start, (int)n_q, (int)dim);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mb_dec_assemble_bwd(const float* d_queries, const float* d_context, const int64_t* ids_keep,
                        const int64_t* ids_restore, float* d_ctx, float* d_emb, float* d_mask_token,
                        void* workspace, int64_t batch, int64_t n_vis, int64_t n_global,
                        int64_t n_all, int64_t q_start, int64_t n_q, int64_t dim, void* stream) {
  MB_REQUIRE(d_queries && d_context && ids_restore && d_ctx && d_emb && d_mask_token && workspace,
             "mb_dec_assemble_bwd: null pointer");
  MB_REQUIRE(dim % 4 == 0, "mb_dec_assemble_bwd: dim must be a multiple of 4");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long rows = batch * (n_vis + n_global);
  if (rows > 0) {
    dec_assemble_bwd_ctx_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(
        d_queries, d_context, reinterpret_cast<const long long*>(ids_keep), d_ctx, (int)batch,
        (int)n_vis, (int)n_global, (int)n_all, (int)q_start, (int)n_q, (int)dim);
    MB_CHECK_CUDA(cudaGetLastError());
  }
  float* part = reinterpret_cast<float*>(workspace);  // [n_q, dim]
  dec_assemble_bwd_emb_kernel<<<(unsigned)n_all, 256, 0, st>>>(
      d_queries, d_context, reinterpret_cast<const long long*>(ids_restore), d_emb, part, (int)batch,
      (int)n_vis, (int)n_global, (int)n_all, (int)q_start, (int)n_q, (int)dim);
  MB_CHECK_CUDA(cudaGetLastError());
  rowsum_small_kernel<<<(unsigned)((dim + 255) / 256), 256, 0, st>>>(part, d_mask_token, (int)n_q,
                                                                     (int)dim);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mb_patchify_cast(const float* img, void* out_bf16, int64_t batch, int32_t channels, int32_t ph,
                     int32_t pw, int32_t gh, int32_t gw, void* stream) {
  MB_REQUIRE(img && out_bf16, "mb_patchify_cast: null pointer");
  MB_REQUIRE(pw % 8 == 0, "mb_patchify_cast: patch width must be a multiple of 8");
  const long long n_vec = batch * gh * gw * (long long)channels * ph * (pw / 8);
  if (n_vec == 0) return 0;
  patchify_cast_kernel<<<(unsigned)((n_vec + 255) / 256), 256, 0,
                         reinterpret_cast<cudaStream_t>(stream)>>>(
      img, reinterpret_cast<__nv_bfloat16*>(out_bf16), n_vec, channels, ph, pw, gh, gw);
  MB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
