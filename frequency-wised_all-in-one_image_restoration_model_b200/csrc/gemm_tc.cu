// K3 - TMA-fed tcgen05 GEMM (fp32 operands read straight from HBM, fp32 accumulators in TMEM) with the same fused
// epilogue as the SIMT kernel.  Covers the dense contractions of the LeWin / LeFF / head path:
//   NT : C[M,N] = A[M,K] . W[N,K]^T           (both operands K-major: every nn.Linear forward)
//   NN : C[M,N] = A[M,K] . B[K,N]             (B MN-major: dX = dY . W)
//   TN : C[M,N] = A[K,M]^T . B[K,N]           (both MN-major: dW = dY^T . X, split over K with fp32 atomics)
//
// Precision.  kind::tf32 TRUNCATES each fp32 operand to a 10-bit mantissa; one pass costs ~1e-3 relative per product
// and pushes the end-to-end restoration error to 4e-3..1.4e-2 (measured), outside north_star's 1e-3.  The default
// (x3) is therefore the error-compensated 3-pass scheme: with hi = trunc_tf32(x) (what the tensor core sees when fed
// x) and lo = x - hi (exact in fp32, <= 13 significant bits),
//      A.B ~= A_lo.B + A.B_lo + A.B          (all three accumulate into the same fp32 TMEM tile)
// where the tensor core again truncates lo to its top 11 bits; the dropped terms are O(2^-21) relative.  lo tiles are
// produced on-chip by 4 "split" warps between TMA arrival and MMA issue (element-wise, so swizzle-agnostic); HBM
// traffic is unchanged, and the layers that dominate the step are HBM-bound.
// The tensor core also accumulates with truncation: the error of a length-k running sum grows ~4e-8*k^1.5 (measured:
// 2.7e-2 at k=8192 on unit-variance data, 20x the fp32 SIMT kernel).  Accumulation in TMEM is therefore limited to
// KC = 256 reduction elements; the epilogue warps promote each partial tile into fp32 registers (round-to-nearest)
// while the MMA fills the other TMEM stage - the same ping-pong that overlaps epilogue and mainloop across tiles.
//
// Persistent, warp-specialised, one CTA per SM: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM alloc), warps 2-5
// split, warps 6-13 epilogue.  smem ring of `stages` x {A 128x32, B BNx32 (+ lo copies)} fp32 tiles in the swizzled
// layouts TMA writes and the UMMA descriptors read.  Operand majors, x3 and the ring depth are runtime flags (they only
// steer single-thread code); BN and the epilogue flavour are template parameters to keep each kernel's code inside
// the instruction cache (a fully generic, fully unrolled epilogue measured ~40 % "no instruction" stalls).
// Every mbarrier wait is bounded: a protocol bug traps instead of hanging the GPU.
#include <cuda.h>
#include <stdlib.h>
#include "freqair_internal.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 32;              // floats per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;           // tf32
constexpr int KC_BLOCKS = 8;        // k-blocks accumulated inside TMEM before promotion to registers (KC = 256)
constexpr int MAX_STAGES = 8;
// 16 warps = 4 per SM sub-partition (128 registers per thread; a 5th warp on a sub-partition would cap everyone at 96):
// warp 0 TMA producer, warp 1 MMA issuer, warps 2-3 B-operand split / rounding, warps 4-7 A-operand staging into TMEM
// (thread = tile row = TMEM lane, quarter = warp % 4), warps 8-15 epilogue.
constexpr int BSPLIT_WARPS = 2;
constexpr int ASPLIT_WARPS = 4;
constexpr int ASPLIT_WARP0 = 2 + BSPLIT_WARPS;
constexpr int SPLIT_WARPS = BSPLIT_WARPS + ASPLIT_WARPS;
constexpr int EPI_WARPS = 8;
constexpr int EPI_WARP0 = 2 + SPLIT_WARPS;     // first epilogue warp (TMEM lane quarter = warp % 4)
constexpr int NTHREADS = 32 * (2 + SPLIT_WARPS + EPI_WARPS);
constexpr int A_BYTES = BM * BK * 4;
// a_tmem: the ring of {A, A_lo} k-blocks (32 or 2 x 32 TMEM columns per stage) starts right behind the two accumulators,
// at column 2 * BN, and takes the rest of the 512 columns: 4 stages for 3x / 2x on 128-wide tiles, 7 on 32-wide ones - the
// huge-M, one-k-block-per-tile contractions of the 128 x 128 level are bound by the LATENCY of the load -> stage -> MMA ->
// free loop (ncu: every role, the TMA warp included, waits ~half of the time with 4 stages in flight), so depth is speed
constexpr int tmem_a_stages(int bn, int passes) { return (512 - 2 * bn) / (passes == 1 ? 32 : 64); }

// epilogue flavours
enum { EPI_PLAIN = 0,   // alpha, accumulate, split-K atomics                         (dX, dW)
       EPI_FWD = 1,     // bias, preact store, activation, row scale, residual        (nn.Linear forward)
       EPI_BWD = 2,     // multiply by act'(aux), accumulate                          (dX through an activation)
       EPI_ANY = 3 };   // every FaGemmEpilogue field (compact, not unrolled)

// per-warp 32x32 fp32 transpose tile of the epilogue, 16-byte chunks XOR-swizzled by row so that both the
// thread=row float4 writes and the lane=column(-quad) reads are bank-conflict-free without padding
__device__ __forceinline__ int stg_idx(int r, int c) { return r * 32 + ((((c >> 2) ^ r) & 7) << 2) + (c & 3); }

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __noinline__ void mbar_timeout() {
  printf("freqair gemm_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) mbar_timeout();     // ~2 s: a pipeline protocol bug, never a legal wait
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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
// 32 lanes x 32 columns of fp32 accumulators -> 32 registers per thread (thread = TMEM lane = output row)
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

// 32 registers per thread -> 32 lanes x 32 columns of TMEM (thread = lane = A-tile row); completion via tc_wait_st()
__device__ __forceinline__ void tc_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),
        "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]),
        "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]),
        "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
      : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),
        "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem: lane = row, column = k] . B[smem descriptor]
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// TS-form MMA with the B descriptor passed as its two 32-bit words (lo carries the 14-bit start-address field, so stepping
// a descriptor is ONE 32-bit add) and the accumulate flag a compile-time constant: the issuing thread is a single lane
// whose own address arithmetic, not the tensor pipe, set the pace of the first version (~65 SASS instructions per
// k-step of 3 MMAs = ~130 cycles per MMA, twice the 64 cycles a 128x128x8 tf32 MMA computes for).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
template <int ACC>
__device__ __forceinline__ void tc_mma_tf32_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
      "mov.b64 bd, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(ACC)
      : "memory");
}

// SS form (both operands from shared memory), descriptors as 32-bit word pairs like tc_mma_tf32_ts2
template <int ACC>
__device__ __forceinline__ void tc_mma_tf32_ss2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 ad, bd;\n\t"
      "mov.b64 ad, {%1, %2};\n\t"
      "mov.b64 bd, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(ACC)
      : "memory");
}

__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// explicit shared-space accesses: the 1024-byte alignment of the dynamic smem base goes through an integer cast, after
// which the compiler only knows a generic pointer and would emit LD.E / ST.E (generic path, long-scoreboard latency)
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}

// round-to-nearest (ties away) to TF32 as the tensor core will see it
// ONE integer add: half an ulp(TF32) is added to the magnitude and the low 13 bits are left for the tensor core to
// truncate (cvt.rna.tf32.f32 compiles to add + inf/nan select + mask: 4 instructions per element on the warps that
// pace the pipeline; the mask is redundant in front of a truncating consumer, the select only matters for inf / nan)
__device__ __forceinline__ float rn1(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
__device__ __forceinline__ float4 rn4(float4 v) { return make_float4(rn1(v.x), rn1(v.y), rn1(v.z), rn1(v.w)); }
__device__ __forceinline__ float lo1(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float4 lo4(float4 v) { return make_float4(lo1(v.x), lo1(v.y), lo1(v.z), lo1(v.w)); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout type
// K-major tile  : box {32 k, rows} -> smem [rows][128 B], TMA swizzle 128B (16-byte chunks XOR row%8), layout 2,
//                 SBO = 1024 B (8 rows), k-step = +32 B inside the swizzle row.
// MN-major tile : box {32 mn, 32 k} per 32-wide MN slab -> smem [slab][32 k][128 B], TMA swizzle 128B_ATOM_32B
//                 (32-byte chunks XOR row%4), layout 1 = SWIZZLE_128B_BASE32B - the ONLY layout the tensor core accepts
//                 for MN-major 32-bit operands (cutlass sm100_common.inl) - LBO = slab stride 4096 B, SBO = atom stride
//                 512 B (4 k-rows), k-step (8 rows) = +1024 B.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, bool mn_major) {
  const uint32_t lbo = mn_major ? 4096u : 16u, sbo = mn_major ? 512u : 1024u, layout = mn_major ? 1u : 2u;
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

struct EpiTC {
  const float* bias; int act; float act_p;
  const float* aux; int64_t ldaux; int aux_act; float aux_p;
  const float* rowscale; int rows_per_scale;
  const float* residual; int64_t ldr;
  int accumulate; float alpha; int atomic;
  float* preact; int64_t ldpre;
  int vec;     // every epilogue array is 16-byte aligned with a row pitch that is a multiple of 4 floats, and N % 4 == 0
  float* a_rowsum;   // a_tmem only: a_rowsum[m] += sum_k op(A)[m,k], accumulated by the split warps (thread = A row)
  const float* a_kscale; int a_krps;   // a_tmem only: op(A)[m,k] *= a_kscale[k / a_krps] (a_krps % 32 == 0: one scalar per k-block)
};

struct TcGeom {
  int M, N, K;
  int64_t ldc;
  int k_chunk, splits, tiles_m, tiles_n;
  int a_mn, b_mn, stages;
  int passes;    // MMAs per product: 3 = A_lo.B + A.B_lo + A.B (fp32-level), 2 = A_lo.rn(B) + A.rn(B) (A exact, B rounded to
                 // nearest TF32), 1 = rn(A).rn(B), 0 = raw operands, truncated by the tensor core (measurement only)
  int a_tmem;    // passes >= 1: A (and A_lo) are staged in TMEM by the split warps (MMA reads only B from shared memory)
  int b_exact;   // passes 1 / 2: B is already TF32-representable (pre-rounded by its producer): the split warps skip it
  int it_stride, it_dsp, it_dtn, it_dtm;     // grid size and its (split, n tile, m tile) decomposition: TileIt::next
  int b_res;     // a_tmem, one n tile, one k-block, no split: every work unit uses the SAME B tile - it is fetched (and split /
                 // rounded) once per ring slot instead of once per unit (148 CTAs re-reading one 3 KB tile made an L2 hot spot:
                 // 20 of the 61 us of 786432 x 28 x 28)
  int a_lin;     // a_tmem, A = [M, K] with lda == K <= 32 (one k-block per tile): the 128 x K tile is a CONTIGUOUS 512 K bytes,
                 // fetched through a flat {32, M K / 32} view of A as 4 K full 128-byte rows instead of 128 rows of 4 K bytes that
                 // straddle 128-byte lines (K = 28: 112 line requests per tile instead of ~240); lands unswizzled, row r at r * 4 K
  // Implicit 3x3 (stride 1, pad 1) patch operand over NHWC tokens x [cB, cH, cW, cC]: the patch matrix
  // col[(b,y,x)][(ky,kx,ci)] = x[b, y+ky-1, x+kx-1, ci] (0 outside the image) is never materialised - the producer
  // addresses x through a 4-D tensor map {C, W, H, B} and lets TMA's out-of-bounds zero fill do the padding.
  //   conv = 1: op(A) = col (K-major; a 128-row M tile is 128/cW image rows of one sample, or a 128-pixel row segment)
  //   conv = 2: op(B) = col (MN-major; N = 9*cC, a 32-wide slab lies inside one tap, a 32-token k-block inside one row)
  int conv, cH, cW, cC;
};

// compile-time activation: keeps the unrolled row loop straight-line (a run-time switch per element splits it into
// basic blocks and serialises the erff chains of the 16 elements a thread has in flight)
template <int ACT>
__device__ __forceinline__ float act_c(float x, float p) {
  if (ACT == ACT_GELU) return gelu_f(x);
  if (ACT == ACT_LRELU) return x > 0.f ? x : x * p;
  if (ACT == ACT_SIGMOID) return 1.0f / (1.0f + __expf(-x));
  return x;
}
template <int ACT>
__device__ __forceinline__ float actg_c(float x, float p) {
  if (ACT == ACT_GELU) return gelu_grad_f(x);
  if (ACT == ACT_LRELU) return x > 0.f ? 1.0f : p;
  if (ACT == ACT_SIGMOID) { const float sg = 1.0f / (1.0f + __expf(-x)); return sg * (1.0f - sg); }
  if (ACT == ACT_MUL) return x;
  return 1.0f;
}

// one float4 (4 consecutive columns of one output row) through the epilogue; ACT = the forward activation (FWD) or
// the activation whose derivative multiplies the result (BWD); EPI_ANY takes both at run time (ACT ignored).
// PRE: the global operands of the row (aux / residual / previous C) were fetched by the caller into e1 / e2, so that
// the loads of a batch of rows are all in flight before the first store (the compiler must assume C aliases them and
// would otherwise serialise one global round trip per row: measured 3.5x on the gelu'(aux) epilogue).
template <int MODE, int ACT, bool PRE>
__device__ __forceinline__ void epi_vec(float4 x, const EpiTC& e, float4 b4, int row, int col, float* __restrict__ cp,
                                        float4 e1, float4 e2) {
  if (MODE == EPI_PLAIN || MODE == EPI_ANY) {
    if (e.atomic) {
      atomicAdd(reinterpret_cast<float4*>(cp), make_float4(x.x * e.alpha, x.y * e.alpha, x.z * e.alpha, x.w * e.alpha));
      return;
    }
  }
  x.x = fmaf(x.x, e.alpha, b4.x); x.y = fmaf(x.y, e.alpha, b4.y);
  x.z = fmaf(x.z, e.alpha, b4.z); x.w = fmaf(x.w, e.alpha, b4.w);
  if (MODE == EPI_FWD || MODE == EPI_ANY) {
    if (e.preact) *reinterpret_cast<float4*>(e.preact + (int64_t)row * e.ldpre + col) = x;
    if (MODE == EPI_ANY) {
      x = make_float4(act_f(x.x, e.act, e.act_p), act_f(x.y, e.act, e.act_p), act_f(x.z, e.act, e.act_p), act_f(x.w, e.act, e.act_p));
    } else if (ACT != ACT_NONE) {
      x = make_float4(act_c<ACT>(x.x, e.act_p), act_c<ACT>(x.y, e.act_p), act_c<ACT>(x.z, e.act_p), act_c<ACT>(x.w, e.act_p));
    }
  }
  if (MODE == EPI_BWD || MODE == EPI_ANY) {
    if (e.aux) {
      const float4 a = PRE ? e1 : __ldg(reinterpret_cast<const float4*>(e.aux + (int64_t)row * e.ldaux + col));
      if (MODE == EPI_ANY) {
        x.x *= act_grad_f(a.x, e.aux_act, e.aux_p); x.y *= act_grad_f(a.y, e.aux_act, e.aux_p);
        x.z *= act_grad_f(a.z, e.aux_act, e.aux_p); x.w *= act_grad_f(a.w, e.aux_act, e.aux_p);
      } else {
        x.x *= actg_c<ACT>(a.x, e.aux_p); x.y *= actg_c<ACT>(a.y, e.aux_p);
        x.z *= actg_c<ACT>(a.z, e.aux_p); x.w *= actg_c<ACT>(a.w, e.aux_p);
      }
    }
    if (MODE == EPI_BWD && e.rowscale) {
      const float rs = __ldg(e.rowscale + row / e.rows_per_scale);
      x.x *= rs; x.y *= rs; x.z *= rs; x.w *= rs;
    }
  }
  if (MODE == EPI_FWD || MODE == EPI_ANY) {
    if (e.rowscale) {
      const float rs = __ldg(e.rowscale + row / e.rows_per_scale);
      x.x *= rs; x.y *= rs; x.z *= rs; x.w *= rs;
    }
    if (e.residual) {
      const float4 a = PRE ? e1 : __ldg(reinterpret_cast<const float4*>(e.residual + (int64_t)row * e.ldr + col));
      x.x += a.x; x.y += a.y; x.z += a.z; x.w += a.w;
    }
  }
  if (MODE != EPI_FWD) {
    if (e.accumulate) {
      const float4 a = PRE ? (MODE == EPI_BWD ? e2 : e1) : *reinterpret_cast<const float4*>(cp);
      x.x += a.x; x.y += a.y; x.z += a.z; x.w += a.w;
    }
  }
  *reinterpret_cast<float4*>(cp) = x;
}

// the 8 row-quads of one 32 x 32 chunk (lane = 4 columns of row it*4 + lane/8), in two batches of four: first every
// global read of the batch, then the arithmetic and the stores
template <int MODE, int ACT>
__device__ __forceinline__ void epi_rows(uint32_t stg_s, const EpiTC& e, float4 b4, int row0, int rsub, int c4, int col,
                                         int M, float* __restrict__ cp0, int64_t ldc) {
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool reads = (MODE == EPI_BWD && (e.aux || e.accumulate)) || (MODE == EPI_FWD && e.residual) ||
                     (MODE == EPI_PLAIN && e.accumulate && !e.atomic);
  if (MODE == EPI_ANY || !reads) {          // store-only epilogue (or the compact generic one): plain row loop
#pragma unroll(MODE == EPI_ANY ? 1 : 4)
    for (int it = 0; it < 8; ++it) {
      const int r = it * 4 + rsub;
      if (row0 + r < M)
        epi_vec<MODE, ACT, false>(lds128(stg_s + 4 * stg_idx(r, c4)), e, b4, row0 + r, col, cp0 + (int64_t)it * 4 * ldc, z4, z4);
    }
    return;
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float4 e1[4], e2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int it = h * 4 + j, row = row0 + it * 4 + rsub;
      e1[j] = z4; e2[j] = z4;
      if (row < M) {
        const float* cp = cp0 + (int64_t)it * 4 * ldc;
        if (MODE == EPI_BWD) {
          if (e.aux) e1[j] = __ldg(reinterpret_cast<const float4*>(e.aux + (int64_t)row * e.ldaux + col));
          if (e.accumulate) e2[j] = *reinterpret_cast<const float4*>(cp);
        } else if (MODE == EPI_FWD) {
          if (e.residual) e1[j] = __ldg(reinterpret_cast<const float4*>(e.residual + (int64_t)row * e.ldr + col));
        } else {
          if (e.accumulate && !e.atomic) e1[j] = *reinterpret_cast<const float4*>(cp);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int it = h * 4 + j, r = it * 4 + rsub;
      if (row0 + r < M)
        epi_vec<MODE, ACT, true>(lds128(stg_s + 4 * stg_idx(r, c4)), e, b4, row0 + r, col, cp0 + (int64_t)it * 4 * ldc, e1[j], e2[j]);
    }
  }
}

// Lean store path for a 32 x 32 chunk that lies completely inside C and whose epilogue only STORES (bias, activation,
// optional pre-activation copy; or the plain alpha * acc): no per-row bound checks, no 64-bit multiplies and no flag tests
// inside the row loop - the pointers advance by one add per row quad and the two swizzled shared-memory offsets (rows
// r = 4*it + rsub alternate between r % 8 = rsub and rsub + 4) are computed once.  ncu on the K = 112, N = 448 contraction
// of the 128 x 128 level: the generic loop executed ~470 instructions per chunk and kept the eight epilogue warps busy
// 90 % of the time while every other role waited; this one is ~4x shorter.
template <int ACT, bool PRE>
__device__ __forceinline__ void epi_chunk_store(uint32_t stg_s, float alpha, float act_p, float4 b4, int rsub, int c4,
                                                float* __restrict__ cp, int64_t ldc, float* __restrict__ pp, int64_t ldpre) {
  const uint32_t a0 = stg_s + 4 * (rsub * 32 + ((((c4 >> 2) ^ rsub) & 7) << 2));
  const uint32_t a1 = stg_s + 4 * ((rsub + 4) * 32 + ((((c4 >> 2) ^ (rsub + 4)) & 7) << 2));
  const int64_t dc = 4 * ldc, dp = 4 * ldpre;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    float4 x = lds128(((it & 1) ? a1 : a0) + (it >> 1) * (8 * 32 * 4));
    x.x = fmaf(x.x, alpha, b4.x); x.y = fmaf(x.y, alpha, b4.y);
    x.z = fmaf(x.z, alpha, b4.z); x.w = fmaf(x.w, alpha, b4.w);
    if (PRE) { *reinterpret_cast<float4*>(pp) = x; pp += dp; }
    if (ACT != ACT_NONE) x = make_float4(act_c<ACT>(x.x, act_p), act_c<ACT>(x.y, act_p), act_c<ACT>(x.z, act_p), act_c<ACT>(x.w, act_p));
    *reinterpret_cast<float4*>(cp) = x;
    cp += dc;
  }
}

// Work units of a persistent CTA: t = blockIdx.x, + gridDim.x, ...; unit t = ((m tile * tiles_n) + n tile) * splits + split.
// Every warp role walks the same sequence; the decomposition is carried incrementally (three adds with carry, the stride's
// own decomposition comes from the host in TcGeom) because two integer divisions per unit and role were a visible part
// of the per-tile hand-off time on the one-k-block shapes.
struct TileIt {
  int t, sp, tn, tm;
  __device__ __forceinline__ TileIt(int t0, const TcGeom& g) {
    t = t0; sp = t0 % g.splits;
    const int r = t0 / g.splits;
    tn = r % g.tiles_n; tm = r / g.tiles_n;
  }
  __device__ __forceinline__ void next(const TcGeom& g) {
    t += g.it_stride;
    sp += g.it_dsp; int c = sp >= g.splits ? 1 : 0; sp -= c ? g.splits : 0;
    tn += g.it_dtn + c; c = tn >= g.tiles_n ? 1 : 0; tn -= c ? g.tiles_n : 0;
    tm += g.it_dtm + c;
  }
};

template <int BN, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB,
                                                              float* __restrict__ C, const TcGeom g, const EpiTC epi) {
  constexpr int B_BYTES = BN * BK * 4;
  constexpr int TMEM_COLS = 2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : 256);
  constexpr int NCHUNK = BN / 32;
  constexpr uint32_t TMEM_A_COL0 = 2 * BN;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int STAGES = g.stages;
  const int PASSES = g.passes;
  const bool X3 = PASSES == 3, SPLIT = PASSES >= 1, A_MN = g.a_mn != 0, B_MN = g.b_mn != 0, A_TMEM = g.a_tmem != 0;
  const uint32_t A_TSTRIDE = PASSES == 1 ? 32u : 64u;      // TMEM columns per stage of the A ring ({A} or {A, A_lo})
  unsigned char* sA = smem;
  unsigned char* sB = sA + STAGES * A_BYTES;
  unsigned char* sAlo = sB + STAGES * B_BYTES;                                // 3 passes only (absent with a_tmem)
  unsigned char* sBlo = sAlo + ((X3 && !A_TMEM) ? STAGES * A_BYTES : 0);
  float* stg_all = reinterpret_cast<float*>(sBlo + (X3 ? STAGES * B_BYTES : 0));   // [EPI_WARPS][32][32]
  uint64_t* full = reinterpret_cast<uint64_t*>(stg_all + EPI_WARPS * 32 * 32);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* ready = empty + MAX_STAGES;       // lo tiles written (x3)
  uint64_t* tmem_full = ready + MAX_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = g.tiles_m * g.tiles_n * g.splits;

  // who publishes a stage: the A warps always; the B warps when they have work (B_lo, or rounding B in place)
  const bool need_b = PASSES == 3 || (SPLIT && !g.b_exact);
  const uint32_t ready_count = A_TMEM ? (ASPLIT_WARPS + (need_b ? BSPLIT_WARPS : 0)) * 32 : SPLIT_WARPS * 32;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], ready_count); }
    // (32-wide N tiles: one 32-column chunk per tile, so the two epilogue warp groups take alternate TILES and each
    // accumulator stage is drained by 4 warps)
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], (NCHUNK == 1 && g.k_chunk <= KC_BLOCKS * BK) ? EPI_WARPS / 2 : EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  const uint32_t tmem_cols = A_TMEM ? 512u : (uint32_t)TMEM_COLS;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    {
      // ---------------------------------------------------------------- TMA producer
      // (whole warp walks the loop, an elected lane issues: coordinates and addresses stay in uniform registers)
      uint32_t ph = 1;               // parity to wait for on empty[s]: flips every time the ring wraps (no division per k-block)
      int s = 0;
      int filled = 0;                // b_res: ring slots whose (single, shared) B tile has been fetched
      for (TileIt it(blockIdx.x, g); it.t < total_tiles; it.next(g)) {
        const int m0 = it.tm * BM, n0 = it.tn * BN;
        const int kbeg = it.sp * g.k_chunk;
        const int nkb = (min(g.K, kbeg + g.k_chunk) - kbeg + BK - 1) / BK;
        for (int kb = 0; kb < nkb; ++kb) {
          const bool load_b = !g.b_res || filled < STAGES;
          filled += (filled < STAGES) ? 1 : 0;
          mbar_wait(&empty[s], ph);
          if (elect_one()) {
          mbar_expect_tx(&full[s], (g.a_lin ? BM * g.K * 4 : A_BYTES) + (load_b ? B_BYTES : 0));
          const int k0 = kbeg + kb * BK;
          if (g.conv == 1) {
            const int hw = g.cH * g.cW, bi = m0 / hw, rem = m0 - bi * hw, y0 = rem / g.cW, x0 = rem - y0 * g.cW;
            const int tap = k0 / g.cC, c0 = k0 - tap * g.cC;
            tma_load_4d(sA + s * A_BYTES, &tmA, &full[s], c0, x0 + tap % 3 - 1, y0 + tap / 3 - 1, bi);
          } else if (g.a_lin) {
            tma_load_2d(sA + s * A_BYTES, &tmA, &full[s], 0, (m0 >> 5) * g.K);
          } else if (!A_MN) {
            tma_load_2d(sA + s * A_BYTES, &tmA, &full[s], k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 32; ++j) tma_load_2d(sA + s * A_BYTES + j * 4096, &tmA, &full[s], m0 + 32 * j, k0);
          }
          if (!load_b) {
          } else if (g.conv == 2) {
            const int hw = g.cH * g.cW, bi = k0 / hw, rem = k0 - bi * hw, y0 = rem / g.cW, x0 = rem - y0 * g.cW;
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) {
              const int n = n0 + 32 * j, tap = n / g.cC, c0 = tap < 9 ? n - tap * g.cC : g.cC;     // past N: channel OOB -> zeros
              tma_load_4d(sB + s * B_BYTES + j * 4096, &tmB, &full[s], c0, x0 + tap % 3 - 1, y0 + tap / 3 - 1, bi);
            }
          } else if (!B_MN) {
            tma_load_2d(sB + s * B_BYTES, &tmB, &full[s], k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) tma_load_2d(sB + s * B_BYTES + j * 4096, &tmB, &full[s], n0 + 32 * j, k0);
          }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ---------------------------------------------------------------- MMA issuer
      // The WHOLE warp walks the loops and waits on the barriers; one elected lane issues.  Under `if (lane == 0)` the
      // compiler cannot keep the descriptors in uniform registers and wraps every tcgen05.mma in an ELECT / R2UR.BROADCAST
      // waterfall (~12 instructions, ~130 cycles per MMA - twice the 64 cycles a 128x128x8 tf32 MMA computes for).
      // instruction descriptor (cute::UMMA::InstrDescriptor): c=f32, a=b=tf32, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (((A_MN && !A_TMEM) ? 1u : 0u) << 15) |
                             ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t astep = A_MN ? 1024u : 32u, bstep = B_MN ? 1024u : 32u;
      uint32_t ph = 0, lu = 0;           // lu = accumulation units issued (a unit = up to KC_BLOCKS k-blocks of one tile)
      int s = 0;
      for (TileIt it(blockIdx.x, g); it.t < total_tiles; it.next(g)) {
        const int kbeg = it.sp * g.k_chunk;
        const int nkb = (min(g.K, kbeg + g.k_chunk) - kbeg + BK - 1) / BK;
        for (int kb0 = 0; kb0 < nkb; kb0 += KC_BLOCKS, ++lu) {
          const uint32_t as = lu & 1;
          mbar_wait(&tmem_empty[as], ((lu >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * BN;
          const int kb1 = min(nkb, kb0 + KC_BLOCKS);
          if (A_TMEM) {
            // descriptors stepped with 32-bit adds: the start-address field (smem address >> 4, < 2^14) never carries
            const uint64_t dz = make_desc(0, B_MN);
            const uint32_t b_hi = (uint32_t)(dz >> 32), lo_c = (uint32_t)dz;
            const uint32_t b_lo0 = lo_c | ((smem_u32(sB) >> 4) & 0x3FFFu), bl_lo0 = lo_c | ((smem_u32(sBlo) >> 4) & 0x3FFFu);
            const uint32_t kinc = bstep >> 4;
            const uint32_t at_base = tmem_base + TMEM_A_COL0;
            for (int kb = kb0; kb < kb1; ++kb) {
              mbar_wait(&ready[s], ph);
              tc_fence_after();
              const uint32_t bs = b_lo0 + (uint32_t)s * (B_BYTES >> 4), bls = bl_lo0 + (uint32_t)s * (B_BYTES >> 4);
              const uint32_t at = at_base + (uint32_t)s * A_TSTRIDE;                         // hi at +0, lo at +32
              if (elect_one()) {
                if (PASSES == 3) {
                  if (kb == kb0) tc_mma_tf32_ts2<0>(d_tmem, at + 32, bs, b_hi, idesc);          // small terms first
                  else tc_mma_tf32_ts2<1>(d_tmem, at + 32, bs, b_hi, idesc);
                  tc_mma_tf32_ts2<1>(d_tmem, at, bls, b_hi, idesc);
                  tc_mma_tf32_ts2<1>(d_tmem, at, bs, b_hi, idesc);
#pragma unroll
                  for (int k = 1; k < BK / UMMA_K; ++k) {
                    tc_mma_tf32_ts2<1>(d_tmem, at + 32 + k * UMMA_K, bs + k * kinc, b_hi, idesc);
                    tc_mma_tf32_ts2<1>(d_tmem, at + k * UMMA_K, bls + k * kinc, b_hi, idesc);
                    tc_mma_tf32_ts2<1>(d_tmem, at + k * UMMA_K, bs + k * kinc, b_hi, idesc);
                  }
                } else if (PASSES == 2) {                  // B was rounded to TF32 in place: A_lo.B + A.B
                  if (kb == kb0) tc_mma_tf32_ts2<0>(d_tmem, at + 32, bs, b_hi, idesc);
                  else tc_mma_tf32_ts2<1>(d_tmem, at + 32, bs, b_hi, idesc);
                  tc_mma_tf32_ts2<1>(d_tmem, at, bs, b_hi, idesc);
#pragma unroll
                  for (int k = 1; k < BK / UMMA_K; ++k) {
                    tc_mma_tf32_ts2<1>(d_tmem, at + 32 + k * UMMA_K, bs + k * kinc, b_hi, idesc);
                    tc_mma_tf32_ts2<1>(d_tmem, at + k * UMMA_K, bs + k * kinc, b_hi, idesc);
                  }
                } else {                                   // both operands rounded to TF32: one MMA per k-step
                  if (kb == kb0) tc_mma_tf32_ts2<0>(d_tmem, at, bs, b_hi, idesc);
                  else tc_mma_tf32_ts2<1>(d_tmem, at, bs, b_hi, idesc);
#pragma unroll
                  for (int k = 1; k < BK / UMMA_K; ++k) tc_mma_tf32_ts2<1>(d_tmem, at + k * UMMA_K, bs + k * kinc, b_hi, idesc);
                }
                tc_commit(&empty[s]);                  // frees the smem slot when these MMAs retire
              }
              __syncwarp();
              if (++s == STAGES) { s = 0; ph ^= 1; }
            }
          } else if (!SPLIT) {
            // raw operands straight from the TMA tiles (SS form), one MMA per k-step: the operands are either
            // TF32-representable already (pre-rounded by their producers) or get truncated by the tensor core
            const uint64_t dza = make_desc(0, A_MN), dzb = make_desc(0, B_MN);
            const uint32_t a_hi = (uint32_t)(dza >> 32), b_hi = (uint32_t)(dzb >> 32);
            const uint32_t a_lo0 = (uint32_t)dza | ((smem_u32(sA) >> 4) & 0x3FFFu), b_lo0 = (uint32_t)dzb | ((smem_u32(sB) >> 4) & 0x3FFFu);
            const uint32_t ainc = astep >> 4, binc = bstep >> 4;
            for (int kb = kb0; kb < kb1; ++kb) {
              mbar_wait(&full[s], ph);
              tc_fence_after();
              const uint32_t as_ = a_lo0 + (uint32_t)s * (A_BYTES >> 4), bs = b_lo0 + (uint32_t)s * (B_BYTES >> 4);
              if (elect_one()) {
                if (kb == kb0) tc_mma_tf32_ss2<0>(d_tmem, as_, a_hi, bs, b_hi, idesc);
                else tc_mma_tf32_ss2<1>(d_tmem, as_, a_hi, bs, b_hi, idesc);
#pragma unroll
                for (int k = 1; k < BK / UMMA_K; ++k) tc_mma_tf32_ss2<1>(d_tmem, as_ + k * ainc, a_hi, bs + k * binc, b_hi, idesc);
                tc_commit(&empty[s]);
              }
              __syncwarp();
              if (++s == STAGES) { s = 0; ph ^= 1; }
            }
          } else
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(SPLIT ? &ready[s] : &full[s], ph);
            tc_fence_after();
            if (lane == 0) {
            const uint32_t a0 = smem_u32(sA + s * A_BYTES), b0 = smem_u32(sB + s * B_BYTES);
            const uint32_t al0 = smem_u32(sAlo + s * A_BYTES), bl0 = smem_u32(sBlo + s * B_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t ad = make_desc(a0 + k * astep, A_MN), bd = make_desc(b0 + k * bstep, B_MN);
              const uint32_t acc0 = (kb > kb0 || k > 0) ? 1u : 0u;
              if (A_TMEM) {
                const uint32_t at = tmem_base + TMEM_A_COL0 + s * 64 + k * UMMA_K;         // hi at +0, lo at +32
                tc_mma_tf32_ts(d_tmem, at + 32, bd, idesc, acc0);                           // small terms first
                tc_mma_tf32_ts(d_tmem, at, make_desc(bl0 + k * bstep, B_MN), idesc, 1u);
                tc_mma_tf32_ts(d_tmem, at, bd, idesc, 1u);
              } else if (X3) {
                tc_mma_tf32(d_tmem, make_desc(al0 + k * astep, A_MN), bd, idesc, acc0);     // small terms first
                tc_mma_tf32(d_tmem, ad, make_desc(bl0 + k * bstep, B_MN), idesc, 1u);
                tc_mma_tf32(d_tmem, ad, bd, idesc, 1u);
              } else {
                tc_mma_tf32(d_tmem, ad, bd, idesc, acc0);
              }
            }
            tc_commit(&empty[s]);                      // frees the smem slot when these MMAs retire
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
          if (lane == 0) tc_commit(&tmem_full[as]);    // partial accumulator complete
          __syncwarp();
        }
      }
    }
  } else if (warp < EPI_WARP0) {
    // ------------------------------------------------------------------ split warps: lo = x - trunc_tf32(x) / rn_tf32(x)
    // These warps pace the kernel (ncu: the only role that never waits), so their loops hold nothing but the loads, one or
    // two ALU ops per element and the TMEM / smem stores.
    if (SPLIT && A_TMEM && warp >= ASPLIT_WARP0) {
      // ---- A warps: thread = A-tile row (= TMEM lane): gather the row's 32 k-values from the swizzled tile TMA wrote,
      // store them (and their lo parts, or their TF32 rounding) into this stage's TMEM columns; the MMA then reads A
      // without touching shared memory
      const int q4 = warp & 3, r = q4 * 32 + lane;
      const bool has_ks = epi.a_kscale != nullptr, has_rs = epi.a_rowsum != nullptr;
      float rsum = 0.f;                       // a_rowsum: this thread's A row, summed over the k-blocks of the tile
      uint32_t ph = 0;
      int s = 0;
      const uint32_t tbase = tmem_base + ((uint32_t)(q4 * 32) << 16) + TMEM_A_COL0;
      const bool A_LIN = g.a_lin != 0;
      const int kq = g.K >> 2;                  // a_lin: float4s per row (odd: the row stride is conflict-free)
      const uint32_t abase = smem_u32(sA) + (A_LIN ? r * g.K * 4 : (A_MN ? (r >> 5) * 4096 + (r & 7) * 4 : r * 128));
      const int cm = (r & 31) >> 3;
      for (TileIt it(blockIdx.x, g); it.t < total_tiles; it.next(g)) {
        const int kbeg = it.sp * g.k_chunk;
        const int nkb = (min(g.K, kbeg + g.k_chunk) - kbeg + BK - 1) / BK;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[s], ph);
          const uint32_t a0 = abase + s * A_BYTES;
          float av[32];
          if (A_LIN) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 v = c < kq ? lds128(a0 + (c << 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
              av[4 * c] = v.x; av[4 * c + 1] = v.y; av[4 * c + 2] = v.z; av[4 * c + 3] = v.w;
            }
          } else if (!A_MN) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 v = lds128(a0 + (((c ^ r) & 7) << 4));
              av[4 * c] = v.x; av[4 * c + 1] = v.y; av[4 * c + 2] = v.z; av[4 * c + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) av[k] = lds32(a0 + k * 128 + (((cm ^ k) & 3) << 5));
          }
          const uint32_t taddr = tbase + s * A_TSTRIDE;
          if (has_ks) {                               // DropPath scale of the k-block's sample, folded into A
            const float sc = __ldg(epi.a_kscale + (kbeg + kb * BK) / epi.a_krps);
#pragma unroll
            for (int k = 0; k < 32; ++k) av[k] *= sc;
          }
          if (has_rs) {
#pragma unroll
            for (int k = 0; k < 32; ++k) rsum += av[k];
          }
          if (PASSES == 1) {
#pragma unroll
            for (int k = 0; k < 32; ++k) av[k] = rn1(av[k]);
            tc_st32(taddr, av);
          } else {
            tc_st32(taddr, av);
#pragma unroll
            for (int k = 0; k < 32; ++k) av[k] = lo1(av[k]);
            tc_st32(taddr + 32, av);
          }
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(&ready[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (has_rs) {
          const int row = it.tm * BM + r;
          if (it.tn == 0 && row < g.M) atomicAdd(epi.a_rowsum + row, rsum);     // each A tile counted once
          rsum = 0.f;
        }
      }
    } else if (SPLIT && A_TMEM) {
      // ---- B warps: lo = x - trunc(x) into the B_lo tile (3 passes), or x rounded to nearest TF32 in place (1 / 2 passes
      // on a B that its producer did not pre-round).  Element-wise, so swizzle-agnostic.
      if (need_b) {
        const int st = threadIdx.x - 64;        // 0..BSPLIT_WARPS*32-1
        constexpr int NB = B_BYTES / 16 / (BSPLIT_WARPS * 32);
        uint32_t ph = 0;
        int s = 0;
        int filled = 0;              // b_res: ring slots whose B tile is already split / rounded (done once per slot)
        for (TileIt it(blockIdx.x, g); it.t < total_tiles; it.next(g)) {
          const int kbeg = it.sp * g.k_chunk;
          const int nkb = (min(g.K, kbeg + g.k_chunk) - kbeg + BK - 1) / BK;
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&full[s], ph);
            if (g.b_res && filled >= STAGES) {               // resident B: nothing to do but publish the stage
              mbar_arrive(&ready[s]);
              if (++s == STAGES) { s = 0; ph ^= 1; }
              continue;
            }
            ++filled;
            const uint32_t b = smem_u32(sB + s * B_BYTES) + st * 16, bl = smem_u32(sBlo + s * B_BYTES) + st * 16;
            float4 rb[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) rb[i] = lds128(b + i * BSPLIT_WARPS * 512);
            if (PASSES == 3) {
#pragma unroll
              for (int i = 0; i < NB; ++i) sts128(bl + i * BSPLIT_WARPS * 512, lo4(rb[i]));
            } else {
#pragma unroll
              for (int i = 0; i < NB; ++i) sts128(b + i * BSPLIT_WARPS * 512, rn4(rb[i]));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy smem writes -> visible to UMMA
            mbar_arrive(&ready[s]);
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    } else if (SPLIT) {
      // ---- legacy 3-pass form with A and A_lo in shared memory (FREQAIR_GEMM_ATMEM=0; kept for A/B measurements):
      // all 6 warps split both tiles.  All loads first, then all stores: a load-convert-store chain per element would
      // serialise ~32 shared-memory round trips per stage.
      const int st = threadIdx.x - 64;        // 0..SPLIT_WARPS*32-1
      uint32_t ph = 0;
      int s = 0;
      for (TileIt it(blockIdx.x, g); it.t < total_tiles; it.next(g)) {
        const int kbeg = it.sp * g.k_chunk;
        const int nkb = (min(g.K, kbeg + g.k_chunk) - kbeg + BK - 1) / BK;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[s], ph);
          const uint32_t a = smem_u32(sA + s * A_BYTES), al = smem_u32(sAlo + s * A_BYTES);
          const uint32_t b = smem_u32(sB + s * B_BYTES), bl = smem_u32(sBlo + s * B_BYTES);
          for (int i = st; i < A_BYTES / 16; i += SPLIT_WARPS * 32) sts128(al + i * 16, lo4(lds128(a + i * 16)));
          for (int i = st; i < B_BYTES / 16; i += SPLIT_WARPS * 32) sts128(bl + i * 16, lo4(lds128(b + i * 16)));
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(&ready[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: warp w owns TMEM lanes 32*(w%4)..+31
    const int ew = warp - EPI_WARP0;          // 0..EPI_WARPS-1
    const int q = warp & 3;
    const int half = ew >> 2;                 // which interleaved set of 32-column chunks: half, half+2
    float* stg = stg_all + ew * (32 * 32);
    const uint32_t stg_s = smem_u32(stg);
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;        // vector path: a warp instruction covers 4 rows x 32 columns
    // epilogues that read nothing from global memory and write every element exactly once take the lean store path
    const bool store_only = (MODE == EPI_PLAIN && !epi.atomic && !epi.accumulate) ||
                            (MODE == EPI_FWD && !epi.residual && !epi.rowscale);
    // ALT (BN = 32): a tile is one chunk; warp group `half` owns every second tile of this CTA (all accumulation units of
    // it) instead of idling - the skinny contractions of the C = 28 level are paced by this epilogue
    // (only when every tile is ONE accumulation unit, k <= 256: unit u then lives in TMEM stage u % 2 = the owning group,
    // so each group sees every phase of its own tmem_full barrier; a group that skipped phases could not tell parity p
    // of the next phase from parity p of the one before)
    const bool ALT = NCHUNK == 1 && g.k_chunk <= KC_BLOCKS * BK;
    uint32_t lu = 0;
    int tix = 0;
    for (TileIt it(blockIdx.x, g); it.t < total_tiles; it.next(g), ++tix) {
      const int m0 = it.tm * BM, n0 = it.tn * BN;
      const int kbeg = it.sp * g.k_chunk;
      const int nkb = (min(g.K, kbeg + g.k_chunk) - kbeg + BK - 1) / BK;
      if (ALT && (tix & 1) != half) {             // the other group's tile: only keep the unit counter in step
        lu += (uint32_t)((nkb + KC_BLOCKS - 1) / KC_BLOCKS);
        continue;
      }
      const int cbase = ALT ? 0 : half;           // chunks cbase, cbase + 2
      float v[2][32];
      for (int kb0 = 0; kb0 < nkb; kb0 += KC_BLOCKS, ++lu) {
        const uint32_t as = lu & 1;
        mbar_wait(&tmem_full[as], (lu >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + as * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          if (cbase + 2 * ci < NCHUNK) {
            if (kb0 == 0) {
              tc_ld32(taddr + (uint32_t)((cbase + 2 * ci) * 32), v[ci]);
            } else {
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                float tmp[16];
                tc_ld16(taddr + (uint32_t)((cbase + 2 * ci) * 32 + hh * 16), tmp);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[ci][hh * 16 + j] += tmp[j];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[as]);           // this TMEM stage may be overwritten
      }
      const int row0 = m0 + q * 32;
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int c = cbase + 2 * ci;
        if (c < NCHUNK && n0 + c * 32 < g.N) {
          // transpose through smem: thread = row writes its 32 columns (8 x float4, conflict-free by the XOR swizzle)
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            sts128(stg_s + 4 * stg_idx(lane, 4 * jj), make_float4(v[ci][4 * jj], v[ci][4 * jj + 1], v[ci][4 * jj + 2], v[ci][4 * jj + 3]));
          __syncwarp();
          if (epi.vec) {
            const int col = n0 + c * 32 + c4;
            if (col < g.N) {                                    // N % 4 == 0 on this path: all four columns valid
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (MODE == EPI_FWD || MODE == EPI_ANY)
                if (epi.bias) b4 = __ldg(reinterpret_cast<const float4*>(epi.bias + col));
              float* cp0 = C + (int64_t)(row0 + rsub) * g.ldc + col;
              const int act = (MODE == EPI_FWD) ? epi.act : (MODE == EPI_BWD ? epi.aux_act : ACT_NONE);
              bool done = false;
              if (store_only && row0 + 32 <= g.M) {
                if (MODE == EPI_PLAIN) {
                  epi_chunk_store<ACT_NONE, false>(stg_s, epi.alpha, 0.f, b4, rsub, c4, cp0, g.ldc, nullptr, 0);
                  done = true;
                } else if (MODE == EPI_FWD) {
                  float* pp0 = epi.preact ? epi.preact + (int64_t)(row0 + rsub) * epi.ldpre + col : nullptr;
                  done = true;
                  if (pp0 && act == ACT_GELU) epi_chunk_store<ACT_GELU, true>(stg_s, epi.alpha, epi.act_p, b4, rsub, c4, cp0, g.ldc, pp0, epi.ldpre);
                  else if (!pp0 && act == ACT_GELU) epi_chunk_store<ACT_GELU, false>(stg_s, epi.alpha, epi.act_p, b4, rsub, c4, cp0, g.ldc, nullptr, 0);
                  else if (!pp0 && act == ACT_LRELU) epi_chunk_store<ACT_LRELU, false>(stg_s, epi.alpha, epi.act_p, b4, rsub, c4, cp0, g.ldc, nullptr, 0);
                  else if (!pp0 && act == ACT_NONE) epi_chunk_store<ACT_NONE, false>(stg_s, epi.alpha, epi.act_p, b4, rsub, c4, cp0, g.ldc, nullptr, 0);
                  else done = false;                      // rarer combinations: generic loop below
                }
              }
              if (done) {
              } else
              if (act == ACT_GELU) epi_rows<MODE, ACT_GELU>(stg_s, epi, b4, row0, rsub, c4, col, g.M, cp0, g.ldc);
              else if (act == ACT_LRELU) epi_rows<MODE, ACT_LRELU>(stg_s, epi, b4, row0, rsub, c4, col, g.M, cp0, g.ldc);
              else if (act == ACT_SIGMOID) epi_rows<MODE, ACT_SIGMOID>(stg_s, epi, b4, row0, rsub, c4, col, g.M, cp0, g.ldc);
              else if (MODE == EPI_BWD && act == ACT_MUL) epi_rows<MODE, ACT_MUL>(stg_s, epi, b4, row0, rsub, c4, col, g.M, cp0, g.ldc);
              else epi_rows<MODE, ACT_NONE>(stg_s, epi, b4, row0, rsub, c4, col, g.M, cp0, g.ldc);
            }
          } else {
            // scalar fallback (ragged N or unaligned rows): lane = column, every FaGemmEpilogue field honoured
            const int col = n0 + c * 32 + lane;
            const bool colok = col < g.N;
            const float bias_v = (epi.bias && colok) ? epi.bias[col] : 0.f;
            const int nrows = min(32, g.M - row0);
#pragma unroll 1
            for (int r = 0; r < nrows; ++r) {
              const int row = row0 + r;
              float x = lds32(stg_s + 4 * stg_idx(r, lane)) * epi.alpha;
              if (!colok) continue;
              float* cp = C + (int64_t)row * g.ldc + col;
              if (epi.atomic) { atomicAdd(cp, x); continue; }
              x += bias_v;
              if (epi.preact) epi.preact[(int64_t)row * epi.ldpre + col] = x;
              x = act_f(x, epi.act, epi.act_p);
              if (epi.aux) x *= act_grad_f(epi.aux[(int64_t)row * epi.ldaux + col], epi.aux_act, epi.aux_p);
              if (epi.rowscale) x *= epi.rowscale[row / epi.rows_per_scale];
              if (epi.residual) x += epi.residual[(int64_t)row * epi.ldr + col];
              if (epi.accumulate) x += *cp;
              *cp = x;
            }
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
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
bool make_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
              bool mn_major, bool flat = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   flat ? CU_TENSOR_MAP_SWIZZLE_NONE : (mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// NHWC tokens x [B, H, W, C] as a 4-D tensor {C, W, H, B}; box = {32 channels, box_w pixels, box_h rows, 1 sample}
bool make_map_nhwc(CUtensorMap* m, const float* base, int B, int H, int W, int C, int box_w, int box_h, bool mn_major) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, int MODE>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, float* C, TcGeom g, const EpiTC& e, cudaStream_t st) {
  // one persistent CTA per SM; the operand ring takes what the 227 KB leave after the 32 KB epilogue transpose tiles
  const int stage_bytes = g.a_tmem ? (A_BYTES + (g.passes == 3 ? 2 : 1) * BN * BK * 4) : (A_BYTES + BN * BK * 4) * (g.passes == 3 ? 2 : 1);
  const int tail_bytes = EPI_WARPS * 32 * 32 * 4 + (3 * MAX_STAGES + 4) * 8 + 16;
  const int budget = 227 * 1024 - 1024 - tail_bytes;
  int stages = budget / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (g.a_tmem && stages > tmem_a_stages(BN, g.passes)) stages = tmem_a_stages(BN, g.passes);
  g.stages = stages;
  const size_t smem = 1024 + (size_t)stages * stage_bytes + tail_bytes;
  auto kern = gemm_tc_kernel<BN, MODE>;
  FA_SMEM_ATTR_ONCE(227 * 1024, kern);
  const int64_t total = (int64_t)g.tiles_m * g.tiles_n * g.splits;
  const int grid = (int)(total < kNumSMs ? total : kNumSMs);
  g.it_stride = grid; g.it_dsp = grid % g.splits;
  g.it_dtn = (grid / g.splits) % g.tiles_n; g.it_dtm = (grid / g.splits) / g.tiles_n;
  kern<<<grid, NTHREADS, smem, st>>>(ta, tb, C, g, e);
  FA_LAUNCH_CHECK("fa_gemm(tcgen05)");
  return FA_OK;
}

template <int MODE>
int dispatch_bn(int bn, const CUtensorMap& ta, const CUtensorMap& tb, float* C, const TcGeom& g, const EpiTC& e,
                cudaStream_t st) {
  switch (bn) {
    case 32: return launch<32, MODE>(ta, tb, C, g, e, st);
    case 64: return launch<64, MODE>(ta, tb, C, g, e, st);
    case 96: return launch<96, MODE>(ta, tb, C, g, e, st);
    default: return launch<128, MODE>(ta, tb, C, g, e, st);
  }
}

}  // namespace

int fa_gemm_tc_launch(const float* A, const float* B, float* C, int M, int N, int K, int64_t lda, int64_t ldb,
                      int64_t ldc, int transA, int transB, const FaGemmEpilogue* ep, cudaStream_t st, int passes,
                      const FaConvOperand* conv) {
  // conv != NULL: the operand it names (1 = A, 2 = B) is the implicit 3x3 patch matrix of the token tensor passed in
  // its place; the caller has set M / N / K and the layout flags as for the materialised matrix
  if (conv) {
    const bool ok = conv->C % 32 == 0 && (conv->H * conv->W) % 128 == 0 && (conv->W % 128 == 0 || 128 % conv->W == 0) &&
                    passes >= 1 && ((conv->which == 1 && !transA) || (conv->which == 2 && !transB && conv->W % 32 == 0));
    if (!ok) return FA_ERR_UNSUPPORTED;
  }
  // eligibility: 16-byte aligned bases and row pitches (TMA), a tile-sized problem, driver entry point present
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(A) || !al16(B) || (!(conv && conv->which == 1) && (lda % 4)) || (!(conv && conv->which == 2) && (ldb % 4)))
    return FA_ERR_UNSUPPORTED;
  if (M < 8 || N < 8 || K < 8) return FA_ERR_UNSUPPORTED;
  const bool a_mn = transA != 0;          // op(A)[m,k] = A[k*lda+m]: reduction index is the row -> MN-major
  const bool b_mn = transB == 0;          // op(B)[k,n] = B[k*ldb+n]
  if (!get_encode()) return FA_ERR_UNSUPPORTED;

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
    e.a_rowsum = ep->a_rowsum;
    e.a_kscale = ep->a_kscale; e.a_krps = ep->a_k_rows_per_scale;
    if (e.a_kscale && (e.a_krps <= 0 || e.a_krps % BK)) return FA_ERR_UNSUPPORTED;
  }
  e.vec = (N % 4 == 0) && al16(C) && (ldc % 4 == 0) && (!e.bias || al16(e.bias)) &&
          (!e.aux || (al16(e.aux) && e.ldaux % 4 == 0)) && (!e.residual || (al16(e.residual) && e.ldr % 4 == 0)) &&
          (!e.preact || (al16(e.preact) && e.ldpre % 4 == 0));
  // a row scale alone rides the FWD flavour, with act'(aux) the BWD flavour (dX through an activation and a DropPath)
  const bool fwd_like = e.bias || e.act != ACT_NONE || (e.rowscale && !e.aux) || e.residual || e.preact;
  const bool plain = !fwd_like && !e.aux && !e.rowscale;
  int mode;
  if (plain) mode = EPI_PLAIN;
  else if (!e.aux && !e.accumulate) mode = EPI_FWD;
  else if (!fwd_like) mode = EPI_BWD;
  else mode = EPI_ANY;

  // Tiling.  N tile = the narrowest of {32, 64, 96, 128} that covers N, else 128.  When that leaves SMs idle:
  //  - plain epilogues split the reduction (fp32 atomics into C; a non-accumulating call zero-fills C first),
  //    which also bounds the length of each truncating tensor-core accumulation;
  //  - the others narrow the N tile (MN-major tiles are gathered in 32-wide slabs, so BN stays a multiple of 32).
  TcGeom g;
  memset(&g, 0, sizeof(g));
  g.M = M; g.N = N; g.K = K; g.ldc = ldc;
  g.a_mn = a_mn; g.b_mn = b_mn; g.passes = passes;
  static const bool atmem_env = [] { const char* e = getenv("FREQAIR_GEMM_ATMEM"); return !(e && e[0] == '0'); }();
  g.a_tmem = (passes >= 1 && (atmem_env || passes != 3)) ? 1 : 0;
  g.b_exact = (ep && ep->b_is_tf32 && passes >= 1 && passes <= 2) ? 1 : 0;
  if (conv) { g.conv = conv->which; g.cH = conv->H; g.cW = conv->W; g.cC = conv->C; }
  if ((e.a_rowsum || e.a_kscale) && !g.a_tmem) return FA_ERR_UNSUPPORTED;      // both ride the TMEM staging of A
  g.tiles_m = (M + BM - 1) / BM;
  int bn = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 96 ? 96 : 128));
  int64_t tiles = (int64_t)g.tiles_m * ((N + bn - 1) / bn);
  g.splits = 1;
  g.k_chunk = K;
  bool split_done = false;
  if (plain && e.vec && K >= 512) {
    // Persistent CTAs walk `rounds` work units each, so the cost of a launch is rounds x (k-extent of a unit + the
    // per-unit pipeline fill / epilogue, ~6 k-blocks).  Splitting the reduction changes how tiles x splits units
    // quantise onto 148 SMs (56 tiles x 6 splits = 336 units = 3 rounds of K/6, but x 5 = 280 units = 2 rounds of K/5);
    // pick the split with the least cost.  A split output is reduced with fp32 atomics (zero-filled first unless the
    // call accumulates); each split also bounds the length of one truncating tensor-core accumulation.
    int maxs = K / 256; if (maxs < 1) maxs = 1;
    const int cap = (int)(4 * kNumSMs / tiles) > 64 ? (int)(4 * kNumSMs / tiles) : 64;
    if (maxs > cap) maxs = cap;
    double best = 1e30; int best_kc = K, best_sp = 1;
    for (int sp = 1; sp <= maxs; ++sp) {
      const int kc = ((K + sp - 1) / sp + BK - 1) / BK * BK;
      const int spe = (K + kc - 1) / kc;
      const int64_t units = tiles * spe;
      const int64_t rounds = (units + kNumSMs - 1) / kNumSMs;
      double cost = (double)rounds * (kc + 192);
      if (spe > 1) cost += e.accumulate ? 32 : 96;               // atomic epilogue (+ zero fill)
      if (cost < best - 1e-9) { best = cost; best_kc = kc; best_sp = spe; }
    }
    if (best_sp > 1) {
      g.k_chunk = best_kc;
      g.splits = best_sp;
      e.atomic = 1;
      if (!e.accumulate) FA_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st));
      split_done = true;
    }
  }
  if (!split_done && tiles < kNumSMs && !(plain && e.vec && K >= 1024)) {
    while (bn > 32 && (bn == 96 || (int64_t)g.tiles_m * ((N + bn - 1) / bn) < kNumSMs)) bn = (bn == 96) ? 64 : bn >> 1;
  }
  g.tiles_n = (N + bn - 1) / bn;
  static const bool alin_env = [] { const char* e = getenv("FREQAIR_GEMM_ALIN"); return !(e && e[0] == '0'); }();
  g.a_lin = (alin_env && g.a_tmem && !a_mn && !conv && g.splits == 1 && K <= BK && K % 4 == 0 && ((K >> 2) & 1) && lda == K &&
             ((int64_t)M * K) % 32 == 0) ? 1 : 0;

  g.b_res = (alin_env && g.a_tmem && !conv && g.splits == 1 && g.tiles_n == 1 && K <= BK) ? 1 : 0;

  CUtensorMap ta, tb;
  bool ok;
  if (g.conv == 1) {                  // 128 tokens = 128 / W whole image rows, or a 128-pixel segment of one row
    const int bw = conv->W >= 128 ? 128 : conv->W;
    ok = make_map_nhwc(&ta, A, conv->B, conv->H, conv->W, conv->C, bw, 128 / bw, false);
  } else if (g.a_lin) ok = make_map(&ta, A, (int64_t)M * K / 32, 32, 32, 32, 4 * K, false, true);   // flat view: box {32, 4 K}
  else if (!a_mn) ok = make_map(&ta, A, M, K, lda, BK, BM, false);          // A [M,K]: box {32 k, 128 m}
  else ok = make_map(&ta, A, K, M, lda, 32, BK, true);                // A stored [K,M]: box {32 m, 32 k}
  if (!ok) { fa_set_error("fa_gemm(tcgen05): cuTensorMapEncodeTiled failed for A"); return FA_ERR_CUDA; }
  if (g.conv == 2) ok = make_map_nhwc(&tb, B, conv->B, conv->H, conv->W, conv->C, 32, 1, true);   // 32 tokens of one row
  else if (!b_mn) ok = make_map(&tb, B, N, K, ldb, BK, bn, false);          // W [N,K]: box {32 k, bn n}
  else ok = make_map(&tb, B, K, N, ldb, 32, BK, true);                // B stored [K,N]: box {32 n, 32 k}
  if (!ok) { fa_set_error("fa_gemm(tcgen05): cuTensorMapEncodeTiled failed for B"); return FA_ERR_CUDA; }
  switch (mode) {
    case EPI_PLAIN: return dispatch_bn<EPI_PLAIN>(bn, ta, tb, C, g, e, st);
    case EPI_FWD: return dispatch_bn<EPI_FWD>(bn, ta, tb, C, g, e, st);
    case EPI_BWD: return dispatch_bn<EPI_BWD>(bn, ta, tb, C, g, e, st);
    default: return dispatch_bn<EPI_ANY>(bn, ta, tb, C, g, e, st);
  }
}
