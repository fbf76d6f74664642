// fp32 SIMT GEMM with fused epilogues: the exact-arithmetic contraction used for shapes the
// tcgen05 path does not cover (ragged K, tiny N, split-K weight gradients) and as the in-library
// cross-check of the tensor-core kernels.
//
//   C[M,N] = epi( alpha * op(A)[M,K] * op(B)[K,N] )
//   op(A): transA ? A[k*lda+m] : A[m*lda+k]      op(B): transB ? B[n*ldb+k] : B[k*ldb+n]
//
// Replaces every nn.Linear / F.linear call of the reference hot path (cuBLAS sgemm there), e.g.
// net/decoder_Uformer.py:121-122 (qkv), :294 (proj), net/utils/leff.py:98,114 (LeFF linears), and
// their autograd backward (dX = dY*W, dW = dY^T*X).
#include "common.cuh"
#include "freqair_internal.h"

namespace {

constexpr int BK = 16;
constexpr int NT = 256;

template <int ROWS, bool CONTIG_K, bool VEC>
__device__ __forceinline__ void g2r(const float* __restrict__ P, int64_t ld, int row0, int k0, int rows_total, int K,
                                    float (&reg)[ROWS * BK / NT], int tid) {
  constexpr int PER = ROWS * BK / NT;       // elements per thread (8 for 128, 4 for 64)
  if (CONTIG_K) {
    // element (row, k) at P[row*ld + k]; a thread owns PER/4 float4 along k
#pragma unroll
    for (int v = 0; v < PER / 4; ++v) {
      int f = tid + v * NT;                 // float4 id: ROWS * (BK/4)
      int row = f / (BK / 4), kq = (f % (BK / 4)) * 4;
      int gr = row0 + row, gk = k0 + kq;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < rows_total) {
        const float* src = P + (int64_t)gr * ld + gk;
        if (VEC && gk + 3 < K) {
          val = *reinterpret_cast<const float4*>(src);
        } else {
          if (gk + 0 < K) val.x = src[0];
          if (gk + 1 < K) val.y = src[1];
          if (gk + 2 < K) val.z = src[2];
          if (gk + 3 < K) val.w = src[3];
        }
      }
      reg[v * 4 + 0] = val.x; reg[v * 4 + 1] = val.y; reg[v * 4 + 2] = val.z; reg[v * 4 + 3] = val.w;
    }
  } else {
    // element (row, k) at P[k*ld + row]; a thread owns PER/4 float4 along row
#pragma unroll
    for (int v = 0; v < PER / 4; ++v) {
      int f = tid + v * NT;                 // float4 id: BK * (ROWS/4)
      int k = f / (ROWS / 4), rq = (f % (ROWS / 4)) * 4;
      int gk = k0 + k, gr = row0 + rq;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gk < K) {
        const float* src = P + (int64_t)gk * ld + gr;
        if (VEC && gr + 3 < rows_total) {
          val = *reinterpret_cast<const float4*>(src);
        } else {
          if (gr + 0 < rows_total) val.x = src[0];
          if (gr + 1 < rows_total) val.y = src[1];
          if (gr + 2 < rows_total) val.z = src[2];
          if (gr + 3 < rows_total) val.w = src[3];
        }
      }
      reg[v * 4 + 0] = val.x; reg[v * 4 + 1] = val.y; reg[v * 4 + 2] = val.z; reg[v * 4 + 3] = val.w;
    }
  }
}

template <int ROWS, bool CONTIG_K>
__device__ __forceinline__ void r2s(float (*S)[ROWS + 4], const float (&reg)[ROWS * BK / NT], int tid) {
  constexpr int PER = ROWS * BK / NT;
  if (CONTIG_K) {
#pragma unroll
    for (int v = 0; v < PER / 4; ++v) {
      int f = tid + v * NT;
      int row = f / (BK / 4), kq = (f % (BK / 4)) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) S[kq + i][row] = reg[v * 4 + i];
    }
  } else {
#pragma unroll
    for (int v = 0; v < PER / 4; ++v) {
      int f = tid + v * NT;
      int k = f / (ROWS / 4), rq = (f % (ROWS / 4)) * 4;
      *reinterpret_cast<float4*>(&S[k][rq]) = make_float4(reg[v * 4], reg[v * 4 + 1], reg[v * 4 + 2], reg[v * 4 + 3]);
    }
  }
}

struct EpiDev {
  const float* bias; int act; float act_p;
  const float* aux; int64_t ldaux; int aux_act; float aux_p;
  const float* rowscale; int rows_per_scale;
  const float* residual; int64_t ldr;
  int accumulate; float alpha; int atomic;
  float* preact; int64_t ldpre;
};

__device__ __forceinline__ float epi_apply(const EpiDev& e, float acc, int m, int n, const float* C, int64_t ldc) {
  float v = acc * e.alpha;
  if (e.bias) v += e.bias[n];
  if (e.preact) e.preact[(int64_t)m * e.ldpre + n] = v;
  v = act_f(v, e.act, e.act_p);
  if (e.aux) v *= act_grad_f(e.aux[(int64_t)m * e.ldaux + n], e.aux_act, e.aux_p);
  if (e.rowscale) v *= e.rowscale[m / e.rows_per_scale];
  if (e.residual) v += e.residual[(int64_t)m * e.ldr + n];
  if (e.accumulate) v += C[(int64_t)m * ldc + n];
  return v;
}

// TM x TN micro-tile per thread; for 8-wide tiles the 8 elements are split 4 + 4 across the two
// halves of the block tile so that shared-memory float4 reads of a quarter-warp stay conflict-free.
template <int BM, int BN, int TM, int TN, bool A_CK, bool B_CK, bool VEC>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                       float* __restrict__ C, int M, int N, int K, int64_t lda,
                                                       int64_t ldb, int64_t ldc, int k_chunk, EpiDev epi) {
  static_assert((BM / TM) * (BN / TN) == NT, "tile/thread mismatch");
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_chunk;
  const int kend = min(K, kbeg + k_chunk);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[BM * BK / NT], rb[BN * BK / NT];
  g2r<BM, A_CK, VEC>(A, lda, m0, kbeg, M, kend, ra, tid);
  g2r<BN, B_CK, VEC>(B, ldb, n0, kbeg, N, kend, rb, tid);
  r2s<BM, A_CK>(As[0], ra, tid);
  r2s<BN, B_CK>(Bs[0], rb, tid);
  __syncthreads();

  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = (k0 + BK) < kend;
    if (more) {
      g2r<BM, A_CK, VEC>(A, lda, m0, k0 + BK, M, kend, ra, tid);
      g2r<BN, B_CK, VEC>(B, ldb, n0, k0 + BK, N, kend, rb, tid);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      if (TM == 8) {
        *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[buf][k][BM / 2 + ty * 4]);
      } else {
        *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      }
      if (TN == 8) {
        *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&Bs[buf][k][BN / 2 + tx * 4]);
      } else {
        *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      r2s<BM, A_CK>(As[buf ^ 1], ra, tid);
      r2s<BN, B_CK>(Bs[buf ^ 1], rb, tid);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ((TM == 8) ? ((i < 4) ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4)) : ty * 4 + i);
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < TN / 4; ++jh) {
      const int nb = n0 + ((TN == 8) ? ((jh == 0) ? tx * 4 : BN / 2 + tx * 4) : tx * 4);
      if (epi.atomic) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (nb + j < N) atomicAdd(&C[(int64_t)m * ldc + nb + j], acc[i][jh * 4 + j] * epi.alpha);
      } else {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (nb + j < N) ? epi_apply(epi, acc[i][jh * 4 + j], m, nb + j, C, ldc) : 0.f;
        float* dst = C + (int64_t)m * ldc + nb;
        if (VEC && nb + 3 < N && ((ldc & 3) == 0)) {
          *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (nb + j < N) dst[j] = v[j];
        }
      }
    }
  }
}

template <int BM, int BN, int TM, int TN>
void launch_cfg(const float* A, const float* B, float* C, int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc,
                bool a_ck, bool b_ck, bool vec, int splits, int k_chunk, const EpiDev& e, cudaStream_t st) {
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
#define FA_GEMM_CASE(ACK, BCK, V)                                                                        \
  gemm_simt_kernel<BM, BN, TM, TN, ACK, BCK, V><<<grid, NT, 0, st>>>(A, B, C, M, N, K, lda, ldb, ldc, k_chunk, e)
  if (a_ck && b_ck) { if (vec) FA_GEMM_CASE(true, true, true); else FA_GEMM_CASE(true, true, false); }
  else if (a_ck && !b_ck) { if (vec) FA_GEMM_CASE(true, false, true); else FA_GEMM_CASE(true, false, false); }
  else if (!a_ck && b_ck) { if (vec) FA_GEMM_CASE(false, true, true); else FA_GEMM_CASE(false, true, false); }
  else { if (vec) FA_GEMM_CASE(false, false, true); else FA_GEMM_CASE(false, false, false); }
#undef FA_GEMM_CASE
}

}  // namespace

int fa_gemm_simt_launch(const float* A, const float* B, float* C, int M, int N, int K, int64_t lda, int64_t ldb,
                        int64_t ldc, int transA, int transB, const FaGemmEpilogue* ep, cudaStream_t st) {
  EpiDev e;
  memset(&e, 0, sizeof(e));
  e.alpha = 1.0f;
  if (ep) {
    e.bias = ep->bias; e.act = ep->act; e.act_p = ep->act_param;
    e.aux = ep->aux; e.ldaux = ep->ldaux; e.aux_act = ep->aux_act; e.aux_p = ep->aux_param;
    e.rowscale = ep->rowscale; e.rows_per_scale = ep->rows_per_scale > 0 ? ep->rows_per_scale : 1;
    e.residual = ep->residual; e.ldr = ep->ldr; e.accumulate = ep->accumulate;
    e.alpha = ep->alpha;
    e.preact = ep->preact; e.ldpre = ep->ldpre;
  }
  const bool a_ck = !transA, b_ck = transB != 0;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = al16(A) && al16(B) && al16(C) && (lda % 4 == 0) && (ldb % 4 == 0);
  // split-K when the output tile grid cannot fill the machine (weight gradients: tiny [N,K] output,
  // reduction over every token).  Partial sums are combined with fp32 atomics, so the epilogue must
  // be a pure (accumulating) scale.
  const bool big = (M > 64 && N > 64);
  const int bm = big ? 128 : 64, bn = big ? 128 : 64;
  const int64_t tiles = (int64_t)((M + bm - 1) / bm) * ((N + bn - 1) / bn);
  int splits = 1;
  const bool plain = !e.bias && e.act == ACT_NONE && !e.aux && !e.rowscale && !e.residual && !e.preact;
  if (plain && e.accumulate && tiles < 2 * kNumSMs && K >= 2048) {
    splits = (int)((4 * kNumSMs + tiles - 1) / tiles);
    int maxs = K / 256; if (maxs < 1) maxs = 1;
    if (splits > maxs) splits = maxs;
  }
  int k_chunk = K;
  if (splits > 1) {
    k_chunk = ((K + splits - 1) / splits + BK - 1) / BK * BK;
    splits = (K + k_chunk - 1) / k_chunk;
    e.atomic = 1;
  }
  if (M <= 0 || N <= 0) return FA_OK;
  if (big) launch_cfg<128, 128, 8, 8>(A, B, C, M, N, K, lda, ldb, ldc, a_ck, b_ck, vec, splits, k_chunk, e, st);
  else launch_cfg<64, 64, 4, 4>(A, B, C, M, N, K, lda, ldb, ldc, a_ck, b_ck, vec, splits, k_chunk, e, st);
  FA_LAUNCH_CHECK("fa_gemm(simt)");
  return FA_OK;
}
