// instantiations of the tcgen05 GEMM with the folded-LayerNorm epilogues (see gemm_impl.cuh, mb_gemm_args)
#include "gemm_impl.cuh"

namespace mb200 {

int dispatch_gemm_ln(int bn, bool pair, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p,
                     cudaStream_t stream) {
#define MB_LN_CASE(E)                                                                                   \
  if (epi == E) {                                                                                       \
    if (pair && bn == 256) return launch_gemm_pair<256, MB_MAJOR_K, 0, 2, E>(ta, tb, p, stream);        \
    if (pair && bn == 128) return launch_gemm_pair<128, MB_MAJOR_K, 0, 2, E>(ta, tb, p, stream);        \
    if (!pair && bn == 128) return launch_gemm_single<128, MB_MAJOR_K, 0, 2, E>(ta, tb, p, stream);     \
    if (!pair && bn == 64) return launch_gemm_single<64, MB_MAJOR_K, 0, 2, E>(ta, tb, p, stream);       \
  }
  MB_LN_CASE(EPI_RES_LN)
  MB_LN_CASE(EPI_BF16_LN)
  MB_LN_CASE(EPI_GELU_LN)
#undef MB_LN_CASE
  return 1;
}

}  // namespace mb200
