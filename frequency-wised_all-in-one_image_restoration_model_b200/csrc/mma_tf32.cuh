// Warp-level tensor-core contractions on shared-memory fp32 tiles for the attention kernels (K2 / K2'):
// mma.sync.m16n8k8 (tf32 inputs, fp32 accumulate) with the same error-compensated 3-pass split as the tcgen05 GEMM:
//   hi = x & 0xFFFFE000 (exactly what the tensor core keeps of x),  lo = x - hi (exact),
//   c += a_lo*b_hi + a_hi*b_lo + a_hi*b_hi      -> dropped terms are O(2^-21) relative, fp32 round-off level.
// The 64 x 64 x {28,56} products of one attention window are far too small for a TMA/tcgen05 pipeline (one CTA owns
// one window-head and interleaves the contractions with softmax and the in-smem FFT band filter), so the legacy
// warp-synchronous MMA is the right tool: operands are read straight from the smem tiles the rest of the kernel uses.
#pragma once
#include "common.cuh"

namespace mma32 {

__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_16n8k8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One warp: C[16 x 8*NT] += A[16 x 8*ksteps] * B[8*ksteps x 8*NT].
//   fa(m, k): element of A, m in [0,16)           fb(k, n): element of B, n in [0, 8*NT)
// Fragment ownership (lane = 4*g + t): A rows g, g+8 / cols t, t+4;  B rows t, t+4 / col g;
// C: c[nt][0..1] = (row g, cols 2t, 2t+1 of tile nt), c[nt][2..3] = (row g+8, same cols).
template <int NT, class FA, class FB>
__device__ __forceinline__ void warp_mma(float (&c)[NT][4], int ksteps, FA fa, FB fb, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll 2
  for (int ks = 0; ks < ksteps; ++ks) {
    const int k0 = ks * 8;
    uint32_t ah[4], al[4];
    split(fa(g, k0 + t), ah[0], al[0]);
    split(fa(g + 8, k0 + t), ah[1], al[1]);
    split(fa(g, k0 + t + 4), ah[2], al[2]);
    split(fa(g + 8, k0 + t + 4), ah[3], al[3]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      uint32_t bh0, bl0, bh1, bl1;
      split(fb(k0 + t, nt * 8 + g), bh0, bl0);
      split(fb(k0 + t + 4, nt * 8 + g), bh1, bl1);
      mma_16n8k8(c[nt], al, bh0, bh1);
      mma_16n8k8(c[nt], ah, bl0, bl1);
      mma_16n8k8(c[nt], ah, bh0, bh1);
    }
  }
}

}  // namespace mma32
