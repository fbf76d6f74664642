// Shared pieces of the window-attention kernels (K2 decoder form in win_attn.cu, K2' joint form in joint_attn.cu):
// operand-tile pitches, window gather, and the three tensor-core contractions of an attention window
//   S  = scale * A . B^T (+ bias + shift mask)       tile_abt   (Q.K^T, dO.V^T)
//   O  = P . V                                        tile_pv<.., false>   (P.V, dS.K)
//   O' = P^T . V                                      tile_pv<.., true>    (P^T.dO, dS^T.Q)
// on 64-row fp32 tiles resident in shared memory, all through mma32::warp_mma (3xTF32, fp32-level accuracy).
// 8 warps take part (threads 0..255); warp w owns output rows 16*(w&3)..+15 and one half of the output columns.
#pragma once
#include "mma_tf32.cuh"

namespace attn {

constexpr int WIN = 8;
constexpr int NTOK = 64;

// smem row pitches of the 64 x hd operand tiles: KP = hd rounded up to the MMA k-step (zero-filled); pitch == 4 (mod 32)
// makes the K-contiguous fragment loads (rows g / g+8, cols t / t+4) bank-conflict-free; a tile read with the token
// index as k (forward V) uses pitch == 8 (mod 32) for the same reason.
template <int HD> struct Pitch;
template <> struct Pitch<28> { static constexpr int KP = 32, HS = 36, HSV = 40; };
template <> struct Pitch<56> { static constexpr int KP = 56, HS = 60, HSV = 72; };
template <> struct Pitch<64> { static constexpr int KP = 64, HS = 68, HSV = 72; };

// gather a 64-token window tile: dst[t][0..KP) = src[rows[t]*ld + col0 + d] (zero for d >= HD: k-padding).
// The copies are cp.async (16 B, L2 only): every tile of an item is in flight at once and the caller waits once with
// load_wait() before the barrier that publishes the tiles - the first version (ld.global -> st.shared per tile) spent
// a third of the backward kernel's samples on four serialised global-memory latencies per item.
template <int HD, int HS, int NTHR>
__device__ __forceinline__ void load_tile(float* dst, const float* __restrict__ src, int64_t ld, int col0,
                                          const int* rows, int tid) {
  constexpr int KP = Pitch<HD>::KP;
  constexpr int V4 = KP / 4;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(dst);
  for (int i = tid; i < NTOK * V4; i += NTHR) {
    const int t = i / V4, d = (i % V4) * 4;
    if (d < HD) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + (uint32_t)(t * HS + d) * 4u),
                   "l"(src + (int64_t)rows[t] * ld + col0 + d)
                   : "memory");
    } else {
      *reinterpret_cast<float4*>(dst + t * HS + d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}
__device__ __forceinline__ void load_wait() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ float rel_bias(const float* bias, int i, int j) {
  return bias[((i >> 3) - (j >> 3) + 7) * 15 + ((i & 7) - (j & 7) + 7)];
}

// out[i*ldo + j] = scale * dot(A_i, B_j) (+ bias[rel(i,j)] + (-100 if label[i] != label[j])),  i, j in [0, 64)
template <int HD, bool BIASMASK>
__device__ __forceinline__ void tile_abt(const float* A, int lda, const float* Bm, int ldb, float* out, int ldo,
                                         float scale, const float* bias, const int* label, int tid) {
  if (tid >= 256) return;
  const int w = tid >> 5, lane = tid & 31;
  const int m0 = (w & 3) * 16, n0 = (w >> 2) * 32;
  float c[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
  mma32::warp_mma<4>(c, Pitch<HD>::KP / 8, [&](int m, int k) { return A[(m0 + m) * lda + k]; },
                     [&](int k, int n) { return Bm[(n0 + n) * ldb + k]; }, lane);
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = m0 + g + ((r & 2) ? 8 : 0), j = n0 + nt * 8 + 2 * t + (r & 1);
      float sv = c[nt][r] * scale;
      if (BIASMASK) {
        sv += rel_bias(bias, i, j);
        if (label[i] != label[j]) sv += -100.0f;
      }
      out[i * ldo + j] = sv;
    }
}

// row softmax over NCOL (64 or 128) columns of the 64 x NCOL tile P (pitch ldp)
template <int NCOL>
__device__ __forceinline__ void softmax_rows(float* P, int ldp, int tid) {
  if (tid >= 256) return;
  constexpr int NC = NCOL / 32;
  const int w = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    float* row = P + (w * 8 + r) * ldp;
    float v[NC];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) { v[c] = row[lane + 32 * c]; m = fmaxf(m, v[c]); }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) { v[c] = expf(v[c] - m); sum += v[c]; }
    const float inv = 1.0f / warp_sum(sum);
#pragma unroll
    for (int c = 0; c < NC; ++c) row[lane + 32 * c] = v[c] * inv;
  }
}

// out[rows[i]*ld + col0 + d] (=, or += when ATOMIC) scale * sum_k P(i,k) * V[k*ldv + d],  i in [0,64), d in [0,HD),
// k in [0,KTOT);  P(i,k) = P[i*ldp + k]  (TRANS: P[k*ldp + i], the transposed tile).
template <int HD, bool TRANS, int KTOT, bool ATOMIC>
__device__ __forceinline__ void tile_pv(const float* P, int ldp, const float* V, int ldv, float* __restrict__ out,
                                        int64_t ld, int col0, const int* rows, float scale, int tid) {
  if (tid >= 256) return;
  constexpr int NTT = (HD + 7) / 8;          // 8-wide output tiles
  constexpr int NTW = (NTT + 1) / 2;         // per warp: tiles (w>>2), (w>>2)+2, ...
  const int w = tid >> 5, lane = tid & 31;
  const int m0 = (w & 3) * 16, nh = w >> 2;
  float c[NTW][4];
#pragma unroll
  for (int i = 0; i < NTW; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
  mma32::warp_mma<NTW>(c, KTOT / 8,
                       [&](int m, int k) { return TRANS ? P[k * ldp + m0 + m] : P[(m0 + m) * ldp + k]; },
                       [&](int k, int n) {
                         const int d = (nh + 2 * (n >> 3)) * 8 + (n & 7);
                         return d < Pitch<HD>::KP ? V[k * ldv + d] : 0.f;
                       },
                       lane);
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < NTW; ++nt) {
    const int d = (nh + 2 * nt) * 8 + 2 * t;
    if (d < HD) {                              // HD is even: both columns of the pair are valid
      float2* p0 = reinterpret_cast<float2*>(out + (int64_t)rows[m0 + g] * ld + col0 + d);
      float2* p1 = reinterpret_cast<float2*>(out + (int64_t)rows[m0 + g + 8] * ld + col0 + d);
      const float2 v0 = make_float2(c[nt][0] * scale, c[nt][1] * scale), v1 = make_float2(c[nt][2] * scale, c[nt][3] * scale);
      if (ATOMIC) { atomicAdd(p0, v0); atomicAdd(p1, v1); }
      else { *p0 = v0; *p1 = v1; }
    }
  }
}

}  // namespace attn
