// generated dispatch table: explicit instantiations of the tcgen05 GEMM (see gemm_impl.cuh)
#include "gemm_impl.cuh"

namespace mb200 {

int dispatch_gemm_single_a(int bn, int layout, int epi, const CUtensorMap& ta, const CUtensorMap& tb,
                const GemmDev& p, cudaStream_t stream) {
  if (layout == LAY_KK_BF16 && bn == 128 && epi == 0)
    return launch_gemm_single<128, MB_MAJOR_K, 0, 2, 0>(ta, tb, p, stream);
  if (layout == LAY_KK_BF16 && bn == 128 && epi == 1)
    return launch_gemm_single<128, MB_MAJOR_K, 0, 2, 1>(ta, tb, p, stream);
  if (layout == LAY_KK_BF16 && bn == 128 && epi == 2)
    return launch_gemm_single<128, MB_MAJOR_K, 0, 2, 2>(ta, tb, p, stream);
  if (layout == LAY_KK_BF16 && bn == 128 && epi == 4)
    return launch_gemm_single<128, MB_MAJOR_K, 0, 2, 4>(ta, tb, p, stream);
  if (layout == LAY_KK_BF16 && bn == 128 && epi == 5)
    return launch_gemm_single<128, MB_MAJOR_K, 0, 2, 5>(ta, tb, p, stream);
  if (layout == LAY_KK_BF16 && bn == 64 && epi == 0)
    return launch_gemm_single<64, MB_MAJOR_K, 0, 2, 0>(ta, tb, p, stream);
  if (layout == LAY_KK_BF16 && bn == 64 && epi == 1)
    return launch_gemm_single<64, MB_MAJOR_K, 0, 2, 1>(ta, tb, p, stream);
  if (layout == LAY_KK_BF16 && bn == 64 && epi == 2)
    return launch_gemm_single<64, MB_MAJOR_K, 0, 2, 2>(ta, tb, p, stream);
  if (layout == LAY_KK_BF16 && bn == 64 && epi == 4)
    return launch_gemm_single<64, MB_MAJOR_K, 0, 2, 4>(ta, tb, p, stream);
  if (layout == LAY_KK_BF16 && bn == 64 && epi == 5)
    return launch_gemm_single<64, MB_MAJOR_K, 0, 2, 5>(ta, tb, p, stream);
  if (layout == LAY_KK_BF16 && bn == 128 && epi == 6)
    return launch_gemm_single<128, MB_MAJOR_K, 0, 2, 6>(ta, tb, p, stream);
  if (layout == LAY_KK_BF16 && bn == 64 && epi == 6)
    return launch_gemm_single<64, MB_MAJOR_K, 0, 2, 6>(ta, tb, p, stream);
  return 1;  // no specialised instantiation
}

}  // namespace mb200
