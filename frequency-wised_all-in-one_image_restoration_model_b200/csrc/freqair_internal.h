#pragma once
#include "../../include/freqair.h"
#include "common.cuh"

int fa_gemm_simt_launch(const float* A, const float* B, float* C, int M, int N, int K, int64_t lda, int64_t ldb,
                        int64_t ldc, int transA, int transB, const FaGemmEpilogue* ep, cudaStream_t st);
// returns FA_ERR_UNSUPPORTED (without setting the error string) when the shape is not eligible
// the operand `which` (1 = A, 2 = B) of a tcgen05 contraction is the implicit 3x3 s1 p1 patch matrix of NHWC tokens [B,H,W,C]
struct FaConvOperand { int which, B, H, W, C; };
int fa_gemm_tc_launch(const float* A, const float* B, float* C, int M, int N, int K, int64_t lda, int64_t ldb,
                      int64_t ldc, int transA, int transB, const FaGemmEpilogue* ep, cudaStream_t st, int passes,
                      const FaConvOperand* conv = nullptr);
// out[m] += sum_k op(A)[m,k] as a separate pass (fallback of FaGemmEpilogue::a_rowsum); defined in norm.cu
int fa_a_rowsum(const float* A, float* out, int M, int K, int64_t lda, int transA, fa_stream_t stream);
