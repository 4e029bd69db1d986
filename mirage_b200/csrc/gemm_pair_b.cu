// generated dispatch table: explicit instantiations of the tcgen05 GEMM (see gemm_impl.cuh)
#include "gemm_impl.cuh"

namespace mb200 {

int dispatch_gemm_pair_b(int bn, int layout, int epi, const CUtensorMap& ta, const CUtensorMap& tb,
                const GemmDev& p, cudaStream_t stream) {
  if (layout == LAY_KMN_BF16 && bn == 256 && epi == 0)
    return launch_gemm_pair<256, MB_MAJOR_K, 1, 2, 0>(ta, tb, p, stream);
  if (layout == LAY_KMN_BF16 && bn == 256 && epi == 3)
    return launch_gemm_pair<256, MB_MAJOR_K, 1, 2, 3>(ta, tb, p, stream);
  if (layout == LAY_KMN_BF16 && bn == 256 && epi == 5)
    return launch_gemm_pair<256, MB_MAJOR_K, 1, 2, 5>(ta, tb, p, stream);
  if (layout == LAY_KMN_BF16 && bn == 128 && epi == 0)
    return launch_gemm_pair<128, MB_MAJOR_K, 1, 2, 0>(ta, tb, p, stream);
  if (layout == LAY_KMN_BF16 && bn == 128 && epi == 3)
    return launch_gemm_pair<128, MB_MAJOR_K, 1, 2, 3>(ta, tb, p, stream);
  if (layout == LAY_KMN_BF16 && bn == 128 && epi == 5)
    return launch_gemm_pair<128, MB_MAJOR_K, 1, 2, 5>(ta, tb, p, stream);
  if (layout == LAY_MNMN_BF16 && bn == 256 && epi == 4)
    return launch_gemm_pair<256, MB_MAJOR_MN, 1, 2, 4>(ta, tb, p, stream);
  if (layout == LAY_MNMN_BF16 && bn == 256 && epi == 5)
    return launch_gemm_pair<256, MB_MAJOR_MN, 1, 2, 5>(ta, tb, p, stream);
  if (layout == LAY_MNMN_BF16 && bn == 128 && epi == 4)
    return launch_gemm_pair<128, MB_MAJOR_MN, 1, 2, 4>(ta, tb, p, stream);
  if (layout == LAY_MNMN_BF16 && bn == 128 && epi == 5)
    return launch_gemm_pair<128, MB_MAJOR_MN, 1, 2, 5>(ta, tb, p, stream);
  if (layout == LAY_PATCH_TF32 && bn == 256 && epi == 5)
    return launch_gemm_pair<256, MB_A_PATCH32, 0, 4, 5>(ta, tb, p, stream);
  if (layout == LAY_PATCH_TF32 && bn == 128 && epi == 5)
    return launch_gemm_pair<128, MB_A_PATCH32, 0, 4, 5>(ta, tb, p, stream);
  if (layout == LAY_PATCH_TF32 && bn == 256 && epi == 6)
    return launch_gemm_pair<256, MB_A_PATCH32, 0, 4, 6>(ta, tb, p, stream);
  if (layout == LAY_PATCH_TF32 && bn == 128 && epi == 6)
    return launch_gemm_pair<128, MB_A_PATCH32, 0, 4, 6>(ta, tb, p, stream);
  if (layout == LAY_KK_TF32 && bn == 256 && epi == 2)   // visible-token patch embedding: tok += A W^T (fp32, in place)
    return launch_gemm_pair<256, MB_MAJOR_K, 0, 4, 2>(ta, tb, p, stream);
  if (layout == LAY_KK_TF32 && bn == 128 && epi == 2)
    return launch_gemm_pair<128, MB_MAJOR_K, 0, 4, 2>(ta, tb, p, stream);
  if (layout == LAY_KK_TF32 && bn == 256 && epi == 5)
    return launch_gemm_pair<256, MB_MAJOR_K, 0, 4, 5>(ta, tb, p, stream);
  if (layout == LAY_KK_TF32 && bn == 128 && epi == 5)
    return launch_gemm_pair<128, MB_MAJOR_K, 0, 4, 5>(ta, tb, p, stream);
  return 1;  // no specialised instantiation
}

}  // namespace mb200
