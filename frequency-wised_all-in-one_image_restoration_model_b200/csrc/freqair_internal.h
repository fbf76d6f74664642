#pragma once
#include "../../include/freqair.h"
#include "common.cuh"

int fa_gemm_simt_launch(const float* A, const float* B, float* C, int M, int N, int K, int64_t lda, int64_t ldb,
                        int64_t ldc, int transA, int transB, const FaGemmEpilogue* ep, cudaStream_t st);
// returns FA_ERR_UNSUPPORTED (without setting the error string) when the shape is not eligible
int fa_gemm_tc_launch(const float* A, const float* B, float* C, int M, int N, int K, int64_t lda, int64_t ldb,
                      int64_t ldc, int transA, int transB, const FaGemmEpilogue* ep, cudaStream_t st, int passes);
// out[m] += sum_k op(A)[m,k] as a separate pass (fallback of FaGemmEpilogue::a_rowsum); defined in norm.cu
int fa_a_rowsum(const float* A, float* out, int M, int K, int64_t lda, int transA, fa_stream_t stream);
