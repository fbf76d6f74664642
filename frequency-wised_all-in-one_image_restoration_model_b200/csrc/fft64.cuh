// 64x64 real 2-D FFT pieces staged in shared memory, 8 threads per 64-point transform:
// each thread runs a radix-8 butterfly in registers, the 8x8 exchange between the two radix-8
// passes is a warp-shuffle transpose (no shared-memory round trip, no block barrier), and two real
// rows ride one complex transform (z = row_a + i*row_b).  Block barriers are only needed between the
// row pass and the column pass.
//
// Layouts:  real map  P[64][PSTR] floats;  half spectrum  sp[64][SPSTR] float2 (columns 0..32).
// Thread roles: tid/8 = transform id, tid%8 = position inside the transform's 8-lane group.
#pragma once
#include "common.cuh"

namespace fft64 {

constexpr int N = 64;
constexpr int NH = 33;
constexpr int PSTR = 68;     // == 4 (mod 32): conflict-free for the FFT row passes and for mma A-fragment loads
constexpr int SPSTR = 33;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// exp(SIGN * 2*pi*i * k / 64), k in 0..63
template <int SIGN>
__device__ __forceinline__ float2 tw64(int k) {
  float s, c;
  sincospif((float)k * (1.0f / 32.0f), &s, &c);
  return make_float2(c, SIGN * s);
}

// 8-point DFT in registers, natural order in and out. SIGN=-1 forward, +1 inverse (unnormalised).
template <int SIGN>
__device__ __forceinline__ void dft8(float2 (&a)[8]) {
  const float r = 0.70710678118654752440f;
  // stage 1 (DIF): pairs (j, j+4)
  float2 b[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) { b[j] = cadd(a[j], a[j + 4]); b[j + 4] = csub(a[j], a[j + 4]); }
  // twiddles W8^j on the lower half: W8 = exp(SIGN*2*pi*i/8)
  // j=1: (r, SIGN*r); j=2: (0, SIGN*1); j=3: (-r, SIGN*r)
  b[5] = make_float2(r * (b[5].x - SIGN * b[5].y), r * (SIGN * b[5].x + b[5].y));
  b[6] = make_float2(-SIGN * b[6].y, SIGN * b[6].x);
  b[7] = make_float2(r * (-b[7].x - SIGN * b[7].y), r * (SIGN * b[7].x - b[7].y));
  // stage 2: within each half, pairs (j, j+2) with twiddle W4^j (j=0,1): W4 = (0, SIGN)
  float2 c[8];
#pragma unroll
  for (int h = 0; h < 8; h += 4) {
    c[h + 0] = cadd(b[h + 0], b[h + 2]);
    c[h + 1] = cadd(b[h + 1], b[h + 3]);
    c[h + 2] = csub(b[h + 0], b[h + 2]);
    float2 t = csub(b[h + 1], b[h + 3]);
    c[h + 3] = make_float2(-SIGN * t.y, SIGN * t.x);
  }
  // stage 3: pairs (j, j+1); outputs are bit-reversed: position p holds X[bitrev3(p)]
  a[0] = cadd(c[0], c[1]); a[4] = csub(c[0], c[1]);
  a[2] = cadd(c[2], c[3]); a[6] = csub(c[2], c[3]);
  a[1] = cadd(c[4], c[5]); a[5] = csub(c[4], c[5]);
  a[3] = cadd(c[6], c[7]); a[7] = csub(c[6], c[7]);
}

// 8x8 transpose across the 8 lanes of a group: lane l, register j  <->  lane j, register l.
__device__ __forceinline__ void transpose8(float2 (&a)[8], int l) {
#pragma unroll
  for (int s = 1; s < 8; s <<= 1) {
    const bool up = (l & s) != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j & s) continue;
      float2 send = up ? a[j] : a[j + s];
      float2 recv;
      recv.x = __shfl_xor_sync(0xffffffffu, send.x, s);
      recv.y = __shfl_xor_sync(0xffffffffu, send.y, s);
      if (up) a[j] = recv; else a[j + s] = recv;
    }
  }
}

// 64-point FFT over an 8-lane group.  In: lane n2 holds x[8*n1 + n2] in a[n1].
// Out: lane k1 holds X[k1 + 8*k2] in a[k2].  Same (index mod 8 -> lane, index / 8 -> register) layout both sides.
template <int SIGN>
__device__ __forceinline__ void fft64_group(float2 (&a)[8], int l) {
  dft8<SIGN>(a);                               // over n1 -> k1, for this lane's n2 = l
#pragma unroll
  for (int k1 = 1; k1 < 8; ++k1) a[k1] = cmul(a[k1], tw64<SIGN>(l * k1));
  transpose8(a, l);                            // lane k1 now holds, in a[n2], the value for (k1, n2)
  dft8<SIGN>(a);                               // over n2 -> k2
}

// ---- row pass, forward: rows (2f, 2f+1) of P -> sp[2f][0..32], sp[2f+1][0..32]
__device__ __forceinline__ void rows_forward(const float* __restrict__ P, float2* __restrict__ sp, int tid) {
  if (tid >= 256) return;
  const int f = tid >> 3, l = tid & 7;
  float2 a[8];
#pragma unroll
  for (int n1 = 0; n1 < 8; ++n1)
    a[n1] = make_float2(P[(2 * f) * PSTR + 8 * n1 + l], P[(2 * f + 1) * PSTR + 8 * n1 + l]);
  fft64_group<-1>(a, l);
  // partner Z[64-k]: k = l + 8*k2 -> lane (8-l)%8, register (l ? 7-k2 : (8-k2)%8)
  const int src = (tid & ~7) | ((8 - l) & 7);
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) {
    float2 other;
    other.x = __shfl_sync(0xffffffffu, a[7 - k2].x, src);
    other.y = __shfl_sync(0xffffffffu, a[7 - k2].y, src);
    if (l == 0) other = a[(8 - k2) & 7];
    const int k = l + 8 * k2;
    if (k <= 32) {
      const float2 z = a[k2];
      sp[(2 * f) * SPSTR + k] = make_float2(0.5f * (z.x + other.x), 0.5f * (z.y - other.y));
      sp[(2 * f + 1) * SPSTR + k] = make_float2(0.5f * (z.y + other.y), 0.5f * (other.x - z.x));
    }
  }
}

// ---- row pass, inverse: sp rows (2f, 2f+1) -> P rows, scaled by `norm` (1/4096 for a full round trip)
__device__ __forceinline__ void rows_inverse(const float2* __restrict__ sp, float* __restrict__ P, float norm, int tid) {
  if (tid >= 256) return;
  const int f = tid >> 3, l = tid & 7;
  float2 a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = l + 8 * j;
    float2 xa, xb;
    if (k <= 32) {
      xa = sp[(2 * f) * SPSTR + k];
      xb = sp[(2 * f + 1) * SPSTR + k];
    } else {
      xa = sp[(2 * f) * SPSTR + 64 - k];
      xb = sp[(2 * f + 1) * SPSTR + 64 - k];
      xa.y = -xa.y; xb.y = -xb.y;
    }
    a[j] = make_float2(xa.x - xb.y, xa.y + xb.x);        // Xa + i*Xb
  }
  fft64_group<1>(a, l);
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) {
    const int n = l + 8 * k2;
    P[(2 * f) * PSTR + n] = a[k2].x * norm;
    P[(2 * f + 1) * PSTR + n] = a[k2].y * norm;
  }
}

// ---- column pass helpers: 33 columns x 8 lanes = 264 threads, run by 9 whole warps (288 threads).
// load column `col` of sp in the group layout / store it back
__device__ __forceinline__ void col_load(const float2* __restrict__ sp, float2 (&a)[8], int col, int l) {
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = sp[(l + 8 * j) * SPSTR + col];
}
__device__ __forceinline__ void col_store(float2* __restrict__ sp, const float2 (&a)[8], int col, int l) {
#pragma unroll
  for (int j = 0; j < 8; ++j) sp[(l + 8 * j) * SPSTR + col] = a[j];
}

// whole filter on a real map held in P: P <- irfft2(rfft2(P) * gain), gain(u,col) = g0 + coef[band[u][col]].
// Must be called by all threads of a block with blockDim.x >= 288 (9 full warps); contains block barriers.
__device__ __forceinline__ void filter_map(float* __restrict__ P, float2* __restrict__ sp,
                                           const uint8_t* __restrict__ band, const float* __restrict__ coef,
                                           float g0, int tid) {
  rows_forward(P, sp, tid);
  __syncthreads();
  if (tid < 288) {                    // 9 whole warps: 33 columns x 8 lanes, the last 24 lanes shadow column 32
    const int col = min(tid >> 3, 32), l = tid & 7;
    const bool live = (tid >> 3) <= 32;
    float2 a[8];
    col_load(sp, a, col, l);
    fft64_group<-1>(a, l);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float g = g0 + coef[band[(l + 8 * j) * NH + col]];
      a[j].x *= g; a[j].y *= g;
    }
    fft64_group<1>(a, l);
    __syncwarp();                     // shadow lanes read column 32 before its owners overwrite it
    if (live) col_store(sp, a, col, l);
  }
  __syncthreads();
  rows_inverse(sp, P, 1.0f / 4096.0f, tid);
  __syncthreads();
}

}  // namespace fft64
