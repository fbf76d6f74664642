// K2 (encoder form) - joint intra/inter-band window attention (FrequencyWindowAttention,
// encoder_Uformer.py:190-313): the L band copies of one 8x8 window attend to each other as L*64 tokens.
// Every (l1,l2) pair gets its own relative-position-bias table.  The reference ADDS a 0/-100 intra|inter band mask
// (:246-254, :281) instead of -inf; a masked score is >= 100 below an unmasked one of the same row (every row keeps
// its own band's / the other bands' same-position token unmasked), so its softmax weight is <= e^-100+O(10) ~ 1e-40
// relative - below one fp32 ulp of the row sum by 33 orders of magnitude.  The kernel therefore evaluates only the
// unmasked (l1,l2) blocks (intra: 1 of L, inter: L-1 of L) and treats the others as exact zeros; the oracle keeps the
// dense -100 form and tests/test_gpu_ops.py::test_joint_attn compares the two.
// One CTA (256 threads) per (sample, window, head); K and V of all L bands stay resident in shared memory
// while the L query blocks are swept, so q/k/v/o each cross HBM once.
#include "freqair_internal.h"

namespace {

constexpr int WIN = 8;
constexpr int NTOK = 64;
constexpr int MAXL = 3;

struct JGeom { int L, B, H, W, heads, shift, nWy, nWx, kind; };

__device__ __forceinline__ void jtoken(const JGeom& g, int wy, int wx, int p, int& pix, int& label) {
  const int sy = wy * WIN + (p >> 3), sx = wx * WIN + (p & 7);
  int y = sy + g.shift, x = sx + g.shift;
  if (y >= g.H) y -= g.H;
  if (x >= g.W) x -= g.W;
  pix = y * g.W + x;
  const int ry = sy < g.H - WIN ? 0 : (sy < g.H - g.shift ? 1 : 2);
  const int rx = sx < g.W - WIN ? 0 : (sx < g.W - g.shift ? 1 : 2);
  label = g.shift > 0 ? ry * 3 + rx : 0;
}

// (l1,l2) block carries the -100 band mask: intra (kind 0) masks other bands, inter (kind 1) masks the own band
__device__ __forceinline__ bool jmasked(const JGeom& g, int l1, int l2) { return g.kind == 0 ? (l1 != l2) : (l1 == l2); }

template <int HD>
__device__ __forceinline__ void jload(float* dst, const float* __restrict__ src, int64_t ld, int col0, int64_t img_row0,
                                      const int* pix, int tid) {
  constexpr int HS = HD + 1, V4 = HD / 4;
  for (int i = tid; i < NTOK * V4; i += 256) {
    const int t = i / V4, d = (i % V4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + (img_row0 + pix[t]) * ld + col0 + d);
    float* o = dst + t * HS + d;
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
}

// S block [64][L*64] for query band l1: scale*q.k + table[l1*L+l2][rel] + band mask + shift mask
template <int HD>
__device__ __forceinline__ void jscores(const float* Q, const float* K, float* S, int SS, const float* bias,
                                        const int* label, const JGeom& g, int l1, float scale, int tid) {
  constexpr int HS = HD + 1;
  const int ty = tid >> 4, tx = tid & 15;
  for (int l2 = 0; l2 < g.L; ++l2) {
    if (jmasked(g, l1, l2)) {
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) S[(ty + 16 * ii) * SS + l2 * NTOK + tx + 16 * jj] = -INFINITY;
      continue;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* Kl = K + l2 * NTOK * HS;
#pragma unroll 4
    for (int d = 0; d < HD; ++d) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Q[(ty + 16 * i) * HS + d];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Kl[(tx + 16 * j) * HS + d];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    const float* bt = bias + (l1 * g.L + l2) * 225;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int i = ty + 16 * ii, j = tx + 16 * jj;
        float s = acc[ii][jj] * scale + bt[((i >> 3) - (j >> 3) + 7) * 15 + ((i & 7) - (j & 7) + 7)];
        if (label[i] != label[j]) s += -100.0f;
        S[i * SS + l2 * NTOK + j] = s;
      }
  }
}

__device__ __forceinline__ void jsoftmax(float* S, int SS, int ncol, int tid) {
  const int w = tid >> 5, lane = tid & 31;
  for (int r = 0; r < 8; ++r) {
    float* row = S + (w * 8 + r) * SS;
    float v[2 * MAXL];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < 2 * MAXL; ++c) {
      v[c] = (lane + 32 * c < ncol) ? row[lane + 32 * c] : -INFINITY;
      m = fmaxf(m, v[c]);
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 2 * MAXL; ++c) {
      v[c] = (lane + 32 * c < ncol) ? expf(v[c] - m) : 0.f;
      sum += v[c];
    }
    const float inv = 1.0f / warp_sum(sum);
#pragma unroll
    for (int c = 0; c < 2 * MAXL; ++c)
      if (lane + 32 * c < ncol) row[lane + 32 * c] = v[c] * inv;
  }
}

template <int HD>
__global__ void __launch_bounds__(256) joint_fwd_kernel(const float* __restrict__ q, int64_t ldq,
                                                        const float* __restrict__ kv, int64_t ldkv,
                                                        float* __restrict__ o, JGeom g, float scale,
                                                        const float* __restrict__ tables) {
  constexpr int HS = HD + 1;
  extern __shared__ __align__(16) float sm[];
  const int L = g.L, SS = L * NTOK + 1;
  float* K = sm;                         // [L*64][HS]
  float* V = K + L * NTOK * HS;
  float* Q = V + L * NTOK * HS;          // [64][HS]
  float* S = Q + NTOK * HS;              // [64][SS]
  float* bias = S + NTOK * SS;           // [L*L][225]
  int* pix = reinterpret_cast<int*>(bias + L * L * 225);
  int* label = pix + NTOK;
  const int tid = threadIdx.x;
  const int C = g.heads * HD;
  int id = blockIdx.x;
  const int h = id % g.heads; id /= g.heads;
  const int wx = id % g.nWx; id /= g.nWx;
  const int wy = id % g.nWy;
  const int b = id / g.nWy;
  const int64_t HW = (int64_t)g.H * g.W;

  if (tid < NTOK) jtoken(g, wy, wx, tid, pix[tid], label[tid]);
  for (int i = tid; i < L * L * 225; i += 256) bias[i] = tables[(int64_t)i * g.heads + h];
  __syncthreads();
  for (int l = 0; l < L; ++l) {
    const int64_t r0 = ((int64_t)l * g.B + b) * HW;
    jload<HD>(K + l * NTOK * HS, kv, ldkv, h * HD, r0, pix, tid);
    jload<HD>(V + l * NTOK * HS, kv, ldkv, C + h * HD, r0, pix, tid);
  }
  for (int l1 = 0; l1 < L; ++l1) {
    const int64_t r0 = ((int64_t)l1 * g.B + b) * HW;
    __syncthreads();
    jload<HD>(Q, q, ldq, h * HD, r0, pix, tid);
    __syncthreads();
    jscores<HD>(Q, K, S, SS, bias, label, g, l1, scale, tid);
    __syncthreads();
    jsoftmax(S, SS, L * NTOK, tid);
    __syncthreads();
    // O = P.V
    const int ty = tid >> 4, tx = tid & 15;
    constexpr int ND = (HD + 15) / 16;
    float acc[4][ND];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int dd = 0; dd < ND; ++dd) acc[i][dd] = 0.f;
    for (int j = 0; j < L * NTOK; ++j) {
      if (jmasked(g, l1, j >> 6)) { j |= 63; continue; }          // whole band has zero weight
      float p[4], v[ND];
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = S[(ty + 16 * i) * SS + j];
#pragma unroll
      for (int dd = 0; dd < ND; ++dd) { const int d = tx + 16 * dd; v[dd] = d < HD ? V[j * HS + d] : 0.f; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int dd = 0; dd < ND; ++dd) acc[i][dd] = fmaf(p[i], v[dd], acc[i][dd]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int dd = 0; dd < ND; ++dd) {
        const int d = tx + 16 * dd;
        if (d < HD) o[(r0 + pix[ty + 16 * i]) * C + h * HD + d] = acc[i][dd];
      }
  }
}

template <int HD>
__global__ void __launch_bounds__(256) joint_bwd_kernel(const float* __restrict__ q, int64_t ldq,
                                                        const float* __restrict__ kv, int64_t ldkv,
                                                        const float* __restrict__ dout, float* __restrict__ dq,
                                                        float* __restrict__ dkv, JGeom g, float scale,
                                                        const float* __restrict__ tables, float* __restrict__ dtables,
                                                        int total_items) {
  constexpr int HS = HD + 1;
  constexpr int ND = (HD + 15) / 16;
  extern __shared__ __align__(16) float sm[];
  const int L = g.L, SS = L * NTOK + 1, LT = L * NTOK;
  float* K = sm;
  float* V = K + LT * HS;
  float* Q = V + LT * HS;
  float* dO = Q + NTOK * HS;
  float* P = dO + NTOK * HS;             // [64][SS]
  float* X = P + NTOK * SS;              // dP -> dS
  float* bias = X + NTOK * SS;
  float* dbias = bias + L * L * 225;
  int* pix = reinterpret_cast<int*>(dbias + L * L * 225);
  int* label = pix + NTOK;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int C = g.heads * HD;
  const int h = blockIdx.x % g.heads;
  const int64_t HW = (int64_t)g.H * g.W;
  for (int i = tid; i < L * L * 225; i += 256) { bias[i] = tables[(int64_t)i * g.heads + h]; dbias[i] = 0.f; }

  for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
    int id = item / g.heads;
    const int wx = id % g.nWx; id /= g.nWx;
    const int wy = id % g.nWy;
    const int b = id / g.nWy;
    __syncthreads();
    if (tid < NTOK) jtoken(g, wy, wx, tid, pix[tid], label[tid]);
    __syncthreads();
    for (int l = 0; l < L; ++l) {
      const int64_t r0 = ((int64_t)l * g.B + b) * HW;
      jload<HD>(K + l * NTOK * HS, kv, ldkv, h * HD, r0, pix, tid);
      jload<HD>(V + l * NTOK * HS, kv, ldkv, C + h * HD, r0, pix, tid);
    }
    // per-thread accumulators of dK / dV: token j = ty + 16*jj (jj < 4L), feature d = tx + 16*dd
    float aK[4 * MAXL][ND], aV[4 * MAXL][ND];
#pragma unroll
    for (int jj = 0; jj < 4 * MAXL; ++jj)
#pragma unroll
      for (int dd = 0; dd < ND; ++dd) { aK[jj][dd] = 0.f; aV[jj][dd] = 0.f; }

    for (int l1 = 0; l1 < L; ++l1) {
      const int64_t r0 = ((int64_t)l1 * g.B + b) * HW;
      __syncthreads();
      jload<HD>(Q, q, ldq, h * HD, r0, pix, tid);
      jload<HD>(dO, dout, C, h * HD, r0, pix, tid);
      __syncthreads();
      jscores<HD>(Q, K, P, SS, bias, label, g, l1, scale, tid);
      // dP[i][c] = dO_i . V_c
      for (int l2 = 0; l2 < L; ++l2) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        if (jmasked(g, l1, l2)) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) X[(ty + 16 * i) * SS + l2 * NTOK + tx + 16 * j] = 0.f;
          continue;
        }
        const float* Vl = V + l2 * NTOK * HS;
#pragma unroll 4
        for (int d = 0; d < HD; ++d) {
          float a[4], bb[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = dO[(ty + 16 * i) * HS + d];
#pragma unroll
          for (int j = 0; j < 4; ++j) bb[j] = Vl[(tx + 16 * j) * HS + d];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) X[(ty + 16 * i) * SS + l2 * NTOK + tx + 16 * j] = acc[i][j];
      }
      __syncthreads();
      jsoftmax(P, SS, LT, tid);
      __syncthreads();
      // dV += P^T.dO (uses P), then dS = P o (dP - rowsum(dP o P)) in X
#pragma unroll
      for (int jj = 0; jj < 4 * MAXL; ++jj) {
        if (jj >= 4 * L) break;
        if (jmasked(g, l1, jj >> 2)) continue;
        const int j = ty + 16 * jj;
        for (int i = 0; i < NTOK; ++i) {
          const float p = P[i * SS + j];
#pragma unroll
          for (int dd = 0; dd < ND; ++dd) {
            const int d = tx + 16 * dd;
            if (d < HD) aV[jj][dd] = fmaf(p, dO[i * HS + d], aV[jj][dd]);
          }
        }
      }
      {
        const int w = tid >> 5, lane = tid & 31;
        for (int r = 0; r < 8; ++r) {
          const int i = w * 8 + r;
          float dot = 0.f;
          for (int c = lane; c < LT; c += 32) dot += P[i * SS + c] * X[i * SS + c];      // masked blocks: P = 0, X = 0
          dot = warp_sum(dot);
          for (int c = lane; c < LT; c += 32) {
            if (jmasked(g, l1, c >> 6)) continue;                                       // dS stays 0 there
            const float ds = P[i * SS + c] * (X[i * SS + c] - dot);
            X[i * SS + c] = ds;
            if (dtables) {
              const int l2 = c >> 6, j = c & 63;
              atomicAdd(&dbias[(l1 * L + l2) * 225 + ((i >> 3) - (j >> 3) + 7) * 15 + ((i & 7) - (j & 7) + 7)], ds);
            }
          }
        }
      }
      __syncthreads();
      // dQ = scale * dS.K  -> global ; dK += scale * dS^T.Q
      {
        float acc[4][ND];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int dd = 0; dd < ND; ++dd) acc[i][dd] = 0.f;
        for (int c = 0; c < LT; ++c) {
          if (jmasked(g, l1, c >> 6)) { c |= 63; continue; }
          float s4[4], k4[ND];
#pragma unroll
          for (int i = 0; i < 4; ++i) s4[i] = X[(ty + 16 * i) * SS + c];
#pragma unroll
          for (int dd = 0; dd < ND; ++dd) { const int d = tx + 16 * dd; k4[dd] = d < HD ? K[c * HS + d] : 0.f; }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int dd = 0; dd < ND; ++dd) acc[i][dd] = fmaf(s4[i], k4[dd], acc[i][dd]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int dd = 0; dd < ND; ++dd) {
            const int d = tx + 16 * dd;
            if (d < HD) dq[(r0 + pix[ty + 16 * i]) * C + h * HD + d] = acc[i][dd] * scale;
          }
      }
#pragma unroll
      for (int jj = 0; jj < 4 * MAXL; ++jj) {
        if (jj >= 4 * L) break;
        if (jmasked(g, l1, jj >> 2)) continue;
        const int j = ty + 16 * jj;
        for (int i = 0; i < NTOK; ++i) {
          const float ds = X[i * SS + j] * scale;
#pragma unroll
          for (int dd = 0; dd < ND; ++dd) {
            const int d = tx + 16 * dd;
            if (d < HD) aK[jj][dd] = fmaf(ds, Q[i * HS + d], aK[jj][dd]);
          }
        }
      }
    }
#pragma unroll
    for (int jj = 0; jj < 4 * MAXL; ++jj) {
      if (jj >= 4 * L) break;
      const int j = ty + 16 * jj;
      const int l = j >> 6, t = j & 63;
      const int64_t row = ((int64_t)l * g.B + b) * HW + pix[t];
#pragma unroll
      for (int dd = 0; dd < ND; ++dd) {
        const int d = tx + 16 * dd;
        if (d < HD) {
          dkv[row * 2 * C + h * HD + d] = aK[jj][dd];
          dkv[row * 2 * C + C + h * HD + d] = aV[jj][dd];
        }
      }
    }
  }
  __syncthreads();
  if (dtables) for (int i = tid; i < L * L * 225; i += 256) atomicAdd(&dtables[(int64_t)i * g.heads + h], dbias[i]);
}

size_t fwd_smem(int L, int HD) {
  const int HS = HD + 1, SS = L * NTOK + 1;
  return sizeof(float) * ((size_t)2 * L * NTOK * HS + NTOK * HS + NTOK * SS + L * L * 225) + sizeof(int) * 2 * NTOK;
}
size_t bwd_smem(int L, int HD) {
  const int HS = HD + 1, SS = L * NTOK + 1;
  return sizeof(float) * ((size_t)2 * L * NTOK * HS + 2 * NTOK * HS + 2 * NTOK * SS + 2 * L * L * 225) + sizeof(int) * 2 * NTOK;
}

int jcheck(const char* who, int L, int B, int H, int W, int heads, int hd, int shift, int kind, int64_t ldq, int64_t ldkv,
           JGeom& g) {
  FA_REQUIRE(L >= 1 && L <= MAXL, "%s: L=%d unsupported (1..3)", who, L);
  FA_REQUIRE(B > 0 && heads > 0, "%s: empty batch/heads", who);
  FA_REQUIRE(H % WIN == 0 && W % WIN == 0, "%s: H=%d W=%d must be multiples of the 8x8 window", who, H, W);
  FA_REQUIRE(hd == 28 || hd == 56, "%s: head_dim=%d unsupported (28, 56)", who, hd);
  FA_REQUIRE(shift == 0 || (shift == 4 && H > WIN && W > WIN), "%s: shift=%d unsupported", who, shift);
  FA_REQUIRE(kind == 0 || kind == 1, "%s: kind must be 0 (intra) or 1 (inter)", who);
  FA_REQUIRE(kind == 0 || L >= 2, "%s: inter-band attention needs L >= 2", who);
  FA_REQUIRE(ldq % 4 == 0 && ldkv % 4 == 0, "%s: row strides must be multiples of 4 floats", who);
  g.L = L; g.B = B; g.H = H; g.W = W; g.heads = heads; g.shift = shift; g.nWy = H / WIN; g.nWx = W / WIN; g.kind = kind;
  return FA_OK;
}

}  // namespace

extern "C" {

int fa_joint_attn_fwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, float* o, int L, int B, int H, int W,
                      int heads, int hd, int shift, float scale, const float* tables, int kind, fa_stream_t stream) {
  FA_REQUIRE(q && kv && o && tables, "fa_joint_attn_fwd: null pointer");
  JGeom g;
  int rc = jcheck("fa_joint_attn_fwd", L, B, H, W, heads, hd, shift, kind, ldq, ldkv, g);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_JOINT_ATTN, st);
  const size_t smem = fwd_smem(L, hd);
  const int items = B * g.nWy * g.nWx * heads;
  if (hd == 28) {
    FA_CUDA(cudaFuncSetAttribute(joint_fwd_kernel<28>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    joint_fwd_kernel<28><<<items, 256, smem, st>>>(q, ldq, kv, ldkv, o, g, scale, tables);
  } else {
    FA_CUDA(cudaFuncSetAttribute(joint_fwd_kernel<56>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    joint_fwd_kernel<56><<<items, 256, smem, st>>>(q, ldq, kv, ldkv, o, g, scale, tables);
  }
  FA_LAUNCH_CHECK("fa_joint_attn_fwd");
  return FA_OK;
}

int fa_joint_attn_bwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, const float* dout, float* dq,
                      float* dkv, int L, int B, int H, int W, int heads, int hd, int shift, float scale,
                      const float* tables, float* dtables, int kind, fa_stream_t stream) {
  FA_REQUIRE(q && kv && dout && dq && dkv && tables, "fa_joint_attn_bwd: null pointer");
  JGeom g;
  int rc = jcheck("fa_joint_attn_bwd", L, B, H, W, heads, hd, shift, kind, ldq, ldkv, g);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_JOINT_ATTN, st);
  const size_t smem = bwd_smem(L, hd);
  FA_REQUIRE(smem <= 227 * 1024, "fa_joint_attn_bwd: L=%d hd=%d needs %zu B of shared memory", L, hd, smem);
  const int items = B * g.nWy * g.nWx * heads;
  int grid = (2 * kNumSMs / heads) * heads;
  if (grid < heads) grid = heads;
  if (grid > items) grid = items;
  if (hd == 28) {
    FA_CUDA(cudaFuncSetAttribute(joint_bwd_kernel<28>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    joint_bwd_kernel<28><<<grid, 256, smem, st>>>(q, ldq, kv, ldkv, dout, dq, dkv, g, scale, tables, dtables, items);
  } else {
    FA_CUDA(cudaFuncSetAttribute(joint_bwd_kernel<56>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    joint_bwd_kernel<56><<<grid, 256, smem, st>>>(q, ldq, kv, ldkv, dout, dq, dkv, g, scale, tables, dtables, items);
  }
  FA_LAUNCH_CHECK("fa_joint_attn_bwd");
  return FA_OK;
}

}  // extern "C"
