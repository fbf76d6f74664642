// K3 - TMA-fed tcgen05 GEMM (kind::tf32, fp32 operands read straight from HBM, fp32 accumulators in TMEM)
// with the same fused epilogue as the SIMT kernel.  Covers the dense contractions of the LeWin / LeFF /
// head path:
//   NT : C[M,N] = A[M,K] . W[N,K]^T           (both operands K-major: every nn.Linear forward)
//   NN : C[M,N] = A[M,K] . B[K,N]             (B MN-major: dX = dY . W)
//   TN : C[M,N] = A[K,M]^T . B[K,N]           (both MN-major: dW = dY^T . X, split over K with fp32 atomics)
// One CTA = one 128 x BN output tile; 6 warps: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM alloc),
// warps 2-5 epilogue (TMEM -> registers -> global).  smem ring of STAGES x {A 128x32, B BNx32} fp32 tiles in
// the 128-byte-swizzled layout TMA writes and the UMMA descriptors read.  Every mbarrier wait is bounded:
// a protocol bug traps instead of hanging the GPU.
#include <cuda.h>
#include "freqair_internal.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 32;              // floats per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;           // tf32
constexpr int NTHREADS = 192;

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {          // ~2 s at 2 GHz: a pipeline protocol bug, never a legal wait
      printf("freqair gemm_tc: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct EpiTC {
  const float* bias; int act; float act_p;
  const float* aux; int64_t ldaux; int aux_act; float aux_p;
  const float* rowscale; int rows_per_scale;
  const float* residual; int64_t ldr;
  int accumulate; float alpha; int atomic;
  float* preact; int64_t ldpre;
};

// A_MN / B_MN: operand is MN-major in global memory (the reduction index is the ROW of the row-major matrix).
// K-major tile  : box {32 k, rows}  -> smem [rows][128 B], one TMA load, k-step = +32 B inside the swizzle row.
// MN-major tile : box {32 mn, 32 k} per 32-wide MN slab -> smem [slab][32 k][128 B]; k-step (8 rows) = +1024 B,
//                 LBO = slab stride = 4096 B, SBO = 1024 B.
template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NTHREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB,
                                                           float* __restrict__ C, int M, int N, int K, int64_t ldc,
                                                           int k_chunk, EpiTC epi) {
  constexpr int A_BYTES = BM * BK * 4;
  constexpr int B_BYTES = BN * BK * 4;
  constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;
  unsigned char* sB = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_chunk;
  const int kend = min(K, kbeg + k_chunk);
  const int nkb = (kend - kbeg + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
        const int k0 = kbeg + kb * BK;
        if (!A_MN) {
          tma_load_2d(sA + s * A_BYTES, &tmA, &full[s], k0, m0);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 32; ++j) tma_load_2d(sA + s * A_BYTES + j * 4096, &tmA, &full[s], m0 + 32 * j, k0);
        }
        if (!B_MN) {
          tma_load_2d(sB + s * B_BYTES, &tmB, &full[s], k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) tma_load_2d(sB + s * B_BYTES + j * 4096, &tmB, &full[s], n0 + 32 * j, k0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      // instruction descriptor (cute::UMMA::InstrDescriptor): c=f32, a=b=tf32, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA + s * A_BYTES), b0 = smem_u32(sB + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t ad = A_MN ? make_desc(a0 + k * 1024, 4096, 1024) : make_desc(a0 + k * 32, 16, 1024);
          const uint64_t bd = B_MN ? make_desc(b0 + k * 1024, 4096, 1024) : make_desc(b0 + k * 32, 16, 1024);
          tc_mma_tf32(tmem_base, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(&empty[s]);                      // frees the smem slot when these MMAs retire
      }
      tc_commit(tmem_full);                        // accumulator complete
    }
  } else {
    // ------------------------------------------------------------------ epilogue: warps 2..5 own TMEM lanes 32*(warp%4)..
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const bool row_ok = row < M;
    const float rs = (epi.rowscale && row_ok) ? epi.rowscale[row / epi.rows_per_scale] : 1.0f;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (!row_ok || n0 + c0 >= N) continue;
      float* crow = C + (int64_t)row * ldc + n0 + c0;
      const int nvalid = min(32, N - (n0 + c0));
      if (epi.atomic) {
        for (int j = 0; j < nvalid; ++j) atomicAdd(crow + j, v[j] * epi.alpha);
        continue;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < nvalid) {
          const int n = n0 + c0 + j;
          float x = v[j] * epi.alpha;
          if (epi.bias) x += epi.bias[n];
          if (epi.preact) epi.preact[(int64_t)row * epi.ldpre + n] = x;
          x = act_f(x, epi.act, epi.act_p);
          if (epi.aux) x *= act_grad_f(epi.aux[(int64_t)row * epi.ldaux + n], epi.aux_act, epi.aux_p);
          x *= rs;
          if (epi.residual) x += epi.residual[(int64_t)row * epi.ldr + n];
          if (epi.accumulate) x += crow[j];
          v[j] = x;
        }
      }
      if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(crow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
        for (int j = 0; j < nvalid; ++j) crow[j] = v[j];
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// row-major matrix [rows, cols] with row stride ld (floats); box = {box_cols (inner), box_rows}
bool make_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, bool A_MN, bool B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, float* C, int M, int N, int K, int64_t ldc, int splits,
           int k_chunk, const EpiTC& e, cudaStream_t st) {
  constexpr int STAGES = (BN >= 128) ? 3 : 4;
  constexpr size_t SMEM = 1024 + STAGES * (BM * BK * 4 + BN * BK * 4) + (2 * STAGES + 1) * 8 + 16;
  auto kern = gemm_tc_kernel<BN, STAGES, A_MN, B_MN>;
  static bool attr_done = false;
  if (!attr_done) {
    FA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    attr_done = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
  kern<<<grid, NTHREADS, SMEM, st>>>(ta, tb, C, M, N, K, ldc, k_chunk, e);
  FA_LAUNCH_CHECK("fa_gemm(tcgen05)");
  return FA_OK;
}

template <bool A_MN, bool B_MN>
int dispatch_bn(int bn, const CUtensorMap& ta, const CUtensorMap& tb, float* C, int M, int N, int K, int64_t ldc,
                int splits, int k_chunk, const EpiTC& e, cudaStream_t st) {
  switch (bn) {
    case 32: return launch<32, A_MN, B_MN>(ta, tb, C, M, N, K, ldc, splits, k_chunk, e, st);
    case 64: return launch<64, A_MN, B_MN>(ta, tb, C, M, N, K, ldc, splits, k_chunk, e, st);
    case 96: return launch<96, A_MN, B_MN>(ta, tb, C, M, N, K, ldc, splits, k_chunk, e, st);
    default: return launch<128, A_MN, B_MN>(ta, tb, C, M, N, K, ldc, splits, k_chunk, e, st);
  }
}

}  // namespace

int fa_gemm_tc_launch(const float* A, const float* B, float* C, int M, int N, int K, int64_t lda, int64_t ldb,
                      int64_t ldc, int transA, int transB, const FaGemmEpilogue* ep, cudaStream_t st, bool probe_only) {
  // eligibility: 16-byte aligned bases and row pitches (TMA), a tile-sized problem, driver entry point present
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(A) || !al16(B) || (lda % 4) || (ldb % 4)) return FA_ERR_UNSUPPORTED;
  if (M < 64 || N < 16 || K < 8) return FA_ERR_UNSUPPORTED;
  const bool a_mn = transA != 0;          // op(A)[m,k] = A[k*lda+m]: reduction index is the row -> MN-major
  const bool b_mn = transB == 0;          // op(B)[k,n] = B[k*ldb+n]
  if (!get_encode()) return FA_ERR_UNSUPPORTED;
  if (probe_only) return FA_OK;

  EpiTC e;
  memset(&e, 0, sizeof(e));
  e.alpha = 1.0f;
  e.rows_per_scale = 1;
  if (ep) {
    e.bias = ep->bias; e.act = ep->act; e.act_p = ep->act_param;
    e.aux = ep->aux; e.ldaux = ep->ldaux; e.aux_act = ep->aux_act; e.aux_p = ep->aux_param;
    e.rowscale = ep->rowscale; e.rows_per_scale = ep->rows_per_scale > 0 ? ep->rows_per_scale : 1;
    e.residual = ep->residual; e.ldr = ep->ldr; e.accumulate = ep->accumulate; e.alpha = ep->alpha;
    e.preact = ep->preact; e.ldpre = ep->ldpre;
  }
  // MN-major tiles are gathered in 32-wide slabs, so BN is a multiple of 32 there
  int bn = N >= 128 ? 128 : (N > 96 ? 128 : (N > 64 ? 96 : (N > 32 ? 64 : 32)));
  const int64_t tiles = (int64_t)((M + BM - 1) / BM) * ((N + bn - 1) / bn);
  int splits = 1, k_chunk = K;
  const bool plain = !e.bias && e.act == ACT_NONE && !e.aux && !e.rowscale && !e.residual && !e.preact;
  if (plain && e.accumulate && tiles < 2 * kNumSMs && K >= 4096) {
    splits = (int)((3 * kNumSMs + tiles - 1) / tiles);
    int maxs = K / 1024; if (maxs < 1) maxs = 1;
    if (splits > maxs) splits = maxs;
    k_chunk = ((K + splits - 1) / splits + BK - 1) / BK * BK;
    splits = (K + k_chunk - 1) / k_chunk;
    if (splits > 1) e.atomic = 1;
  }
  CUtensorMap ta, tb;
  bool ok;
  if (!a_mn) ok = make_map(&ta, A, M, K, lda, BK, BM);          // A [M,K]: box {32 k, 128 m}
  else ok = make_map(&ta, A, K, M, lda, 32, BK);                // A stored [K,M]: box {32 m, 32 k}
  if (!ok) { fa_set_error("fa_gemm(tcgen05): cuTensorMapEncodeTiled failed for A"); return FA_ERR_CUDA; }
  if (!b_mn) ok = make_map(&tb, B, N, K, ldb, BK, bn);          // W [N,K]: box {32 k, bn n}
  else ok = make_map(&tb, B, K, N, ldb, 32, BK);                // B stored [K,N]: box {32 n, 32 k}
  if (!ok) { fa_set_error("fa_gemm(tcgen05): cuTensorMapEncodeTiled failed for B"); return FA_ERR_CUDA; }
  if (!a_mn && !b_mn) return dispatch_bn<false, false>(bn, ta, tb, C, M, N, K, ldc, splits, k_chunk, e, st);
  if (!a_mn && b_mn) return dispatch_bn<false, true>(bn, ta, tb, C, M, N, K, ldc, splits, k_chunk, e, st);
  if (a_mn && !b_mn) return dispatch_bn<true, false>(bn, ta, tb, C, M, N, K, ldc, splits, k_chunk, e, st);
  return dispatch_bn<true, true>(bn, ta, tb, C, M, N, K, ldc, splits, k_chunk, e, st);
}
