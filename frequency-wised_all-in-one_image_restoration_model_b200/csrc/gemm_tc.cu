// tcgen05 (kind::tf32) + TMA GEMM - placeholder until the tensor-core kernel lands: reports "not eligible"
// so that fa_gemm routes everything to the fp32 SIMT kernel.
#include "freqair_internal.h"

int fa_gemm_tc_launch(const float*, const float*, float*, int, int, int, int64_t, int64_t, int64_t, int, int,
                      const FaGemmEpilogue*, cudaStream_t, bool) {
  return FA_ERR_UNSUPPORTED;
}
