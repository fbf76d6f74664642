// K2' (encoder form) - joint intra/inter-band window attention (FrequencyWindowAttention,
// encoder_Uformer.py:190-313): the L band copies of one 8x8 window attend to each other as L*64 tokens.
// Every (l1,l2) pair gets its own relative-position-bias table.  The reference ADDS a 0/-100 intra|inter band mask
// (:246-254, :281) instead of -inf; a masked score is >= 100 below an unmasked one of the same row (every row keeps
// its own band's / the other bands' same-position token unmasked), so its softmax weight is <= e^-100+O(10) ~ 1e-40
// relative - below one fp32 ulp of the row sum by 33 orders of magnitude.  The kernels therefore evaluate only the
// unmasked (l1,l2) blocks - intra: the query band's own 64 keys, inter: the 64*(L-1) keys of the other bands - and treat
// the rest as exact zeros; the oracle keeps the dense -100 form and tests/test_gpu_ops.py::test_joint_attn compares.
// One CTA (256 threads) per (query band, sample, window, head): the three contractions run on the tensor cores
// (attn_tiles.cuh, 3xTF32), q/k/v/o cross HBM once, windows and the cyclic shift are index arithmetic.
// Backward recomputes P, forms dS = P o (dP - rowsum(P o dP)) from the dP MMA fragments in registers (no second
// 64 x 128 buffer), reduces the bias-table gradients in shared memory across the windows a CTA visits, and adds the
// key-side gradients of a band with fp32 atomics only when two query bands contribute (inter, L = 3; two addends
// commute, so the result stays deterministic).
#include "freqair_internal.h"
#include "attn_tiles.cuh"

namespace {

using attn::WIN;
using attn::NTOK;
constexpr int MAXL = 3;

struct JGeom { int L, B, H, W, heads, shift, nWy, nWx, kind; };

// pixel (y*W + x) of window position p after the cyclic shift, and its SW-MSA region label
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

// key bands of query band l1: intra -> {l1}; inter -> every other band in increasing order
template <int NKB>
__device__ __forceinline__ void key_bands(const JGeom& g, int l1, int (&kb)[NKB]) {
#pragma unroll
  for (int n = 0; n < NKB; ++n) kb[n] = l1;
  if (g.kind == 0) return;
  int n = 0;
  for (int l2 = 0; l2 < g.L; ++l2)
    if (l2 != l1 && n < NKB) kb[n++] = l2;
}

template <int HD, int NKB>
struct JSmemF {
  static constexpr int HS = attn::Pitch<HD>::HS, HSV = attn::Pitch<HD>::HSV, PS = NKB * NTOK + 4;
  float q[NTOK * HS];
  float k[NKB * NTOK * HS];
  float v[NKB * NTOK * HSV];
  float p[NTOK * PS];
  float bias[NKB][232];
  int rowq[NTOK];
  int rowk[NKB][NTOK];
  int label[NTOK];
};

template <int HD, int NKB>
__global__ void __launch_bounds__(256) joint_fwd_kernel(const float* __restrict__ q, int64_t ldq,
                                                        const float* __restrict__ kv, int64_t ldkv,
                                                        float* __restrict__ o, JGeom g, float scale,
                                                        const float* __restrict__ tables) {
  using S = JSmemF<HD, NKB>;
  extern __shared__ __align__(16) unsigned char smraw[];
  S& s = *reinterpret_cast<S*>(smraw);
  const int tid = threadIdx.x;
  const int C = g.heads * HD;
  const int l1 = blockIdx.y;
  int id = blockIdx.x;
  const int h = id % g.heads; id /= g.heads;
  const int wx = id % g.nWx; id /= g.nWx;
  const int wy = id % g.nWy;
  const int b = id / g.nWy;
  const int HW = g.H * g.W;
  int kb[NKB];
  key_bands<NKB>(g, l1, kb);

  if (tid < NTOK) {
    int pix, label;
    jtoken(g, wy, wx, tid, pix, label);
    s.label[tid] = label;
    s.rowq[tid] = (l1 * g.B + b) * HW + pix;
#pragma unroll
    for (int n = 0; n < NKB; ++n) s.rowk[n][tid] = (kb[n] * g.B + b) * HW + pix;
  }
#pragma unroll
  for (int n = 0; n < NKB; ++n)
    for (int i = tid; i < 225; i += 256) s.bias[n][i] = tables[((int64_t)(l1 * g.L + kb[n]) * 225 + i) * g.heads + h];
  __syncthreads();
  attn::load_tile<HD, S::HS, 256>(s.q, q, ldq, h * HD, s.rowq, tid);
#pragma unroll
  for (int n = 0; n < NKB; ++n) {
    attn::load_tile<HD, S::HS, 256>(s.k + n * NTOK * S::HS, kv, ldkv, h * HD, s.rowk[n], tid);
    attn::load_tile<HD, S::HSV, 256>(s.v + n * NTOK * S::HSV, kv, ldkv, C + h * HD, s.rowk[n], tid);
  }
  attn::load_wait();
  __syncthreads();
#pragma unroll
  for (int n = 0; n < NKB; ++n)
    attn::tile_abt<HD, true>(s.q, S::HS, s.k + n * NTOK * S::HS, S::HS, s.p + n * NTOK, S::PS, scale, s.bias[n], s.label, tid);
  __syncthreads();
  attn::softmax_rows<NKB * NTOK>(s.p, S::PS, tid);
  __syncthreads();
  attn::tile_pv<HD, false, NKB * NTOK, false>(s.p, S::PS, s.v, S::HSV, o, C, h * HD, s.rowq, 1.0f, tid);
}

template <int HD, int NKB>
struct JSmemB {
  static constexpr int HS = attn::Pitch<HD>::HS, PS = NKB * NTOK + 4;
  float q[NTOK * HS];
  float dO[NTOK * HS];
  float k[NKB * NTOK * HS];
  float v[NKB * NTOK * HS];
  float p[NTOK * PS];            // S -> P -> dS
  float red[2][NTOK];
  float bias[NKB][232];
  float dbias[NKB][232];
  int rowq[NTOK];
  int rowk[NKB][NTOK];
  int label[NTOK];
};

template <int HD, int NKB, bool ATOMIC>
__global__ void __launch_bounds__(256) joint_bwd_kernel(const float* __restrict__ q, int64_t ldq,
                                                        const float* __restrict__ kv, int64_t ldkv,
                                                        const float* __restrict__ dout, float* __restrict__ dq,
                                                        float* __restrict__ dkv, JGeom g, float scale,
                                                        const float* __restrict__ tables, float* __restrict__ dtables,
                                                        int total_items) {
  using S = JSmemB<HD, NKB>;
  extern __shared__ __align__(16) unsigned char smraw[];
  S& s = *reinterpret_cast<S*>(smraw);
  const int tid = threadIdx.x;
  const int C = g.heads * HD;
  const int l1 = blockIdx.y;
  // gridDim.x is a multiple of heads, so every item this CTA visits has the same head (and the same query band)
  const int h = blockIdx.x % g.heads;
  const int HW = g.H * g.W;
  int kb[NKB];
  key_bands<NKB>(g, l1, kb);
#pragma unroll
  for (int n = 0; n < NKB; ++n)
    for (int i = tid; i < 232; i += 256) {
      s.bias[n][i] = i < 225 ? tables[((int64_t)(l1 * g.L + kb[n]) * 225 + i) * g.heads + h] : 0.f;
      s.dbias[n][i] = 0.f;
    }

  for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
    int id = item / g.heads;
    const int wx = id % g.nWx; id /= g.nWx;
    const int wy = id % g.nWy;
    const int b = id / g.nWy;
    __syncthreads();
    if (tid < NTOK) {
      int pix, label;
      jtoken(g, wy, wx, tid, pix, label);
      s.label[tid] = label;
      s.rowq[tid] = (l1 * g.B + b) * HW + pix;
#pragma unroll
      for (int n = 0; n < NKB; ++n) s.rowk[n][tid] = (kb[n] * g.B + b) * HW + pix;
    }
    __syncthreads();
    attn::load_tile<HD, S::HS, 256>(s.q, q, ldq, h * HD, s.rowq, tid);
    attn::load_tile<HD, S::HS, 256>(s.dO, dout, C, h * HD, s.rowq, tid);
#pragma unroll
    for (int n = 0; n < NKB; ++n) {
      attn::load_tile<HD, S::HS, 256>(s.k + n * NTOK * S::HS, kv, ldkv, h * HD, s.rowk[n], tid);
      attn::load_tile<HD, S::HS, 256>(s.v + n * NTOK * S::HS, kv, ldkv, C + h * HD, s.rowk[n], tid);
    }
    attn::load_wait();
    __syncthreads();
#pragma unroll
    for (int n = 0; n < NKB; ++n)
      attn::tile_abt<HD, true>(s.q, S::HS, s.k + n * NTOK * S::HS, S::HS, s.p + n * NTOK, S::PS, scale, s.bias[n], s.label, tid);
    __syncthreads();
    attn::softmax_rows<NKB * NTOK>(s.p, S::PS, tid);
    __syncthreads();
    // dV[kb] = P[:, kb]^T . dO
#pragma unroll
    for (int n = 0; n < NKB; ++n)
      attn::tile_pv<HD, true, NTOK, ATOMIC>(s.p + n * NTOK, S::PS, s.dO, S::HS, dkv, 2 * C, C + h * HD, s.rowk[n], 1.0f, tid);
    // dP = dO . V^T as MMA fragments (warp w: rows 16*(w&3).., column half w>>2 of every key band)
    const int w = tid >> 5, lane = tid & 31;
    const int m0 = (w & 3) * 16, n0 = (w >> 2) * 32;
    const int gq = lane >> 2, tq = lane & 3;
    float dp[NKB][4][4];
    float part0 = 0.f, part1 = 0.f;                 // sum_j P*dP over this thread's columns, rows m0+gq and m0+gq+8
#pragma unroll
    for (int n = 0; n < NKB; ++n) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { dp[n][i][0] = dp[n][i][1] = dp[n][i][2] = dp[n][i][3] = 0.f; }
      const float* Vn = s.v + n * NTOK * S::HS;
      mma32::warp_mma<4>(dp[n], attn::Pitch<HD>::KP / 8, [&](int m, int k) { return s.dO[(m0 + m) * S::HS + k]; },
                         [&](int k, int nn) { return Vn[(n0 + nn) * S::HS + k]; }, lane);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int i = m0 + gq + ((r & 2) ? 8 : 0), j = n * NTOK + n0 + nt * 8 + 2 * tq + (r & 1);
          const float pv = s.p[i * S::PS + j] * dp[n][nt][r];
          if (r & 2) part1 += pv; else part0 += pv;
        }
    }
    part0 += __shfl_xor_sync(0xffffffffu, part0, 1); part0 += __shfl_xor_sync(0xffffffffu, part0, 2);
    part1 += __shfl_xor_sync(0xffffffffu, part1, 1); part1 += __shfl_xor_sync(0xffffffffu, part1, 2);
    if (tq == 0) { s.red[w >> 2][m0 + gq] = part0; s.red[w >> 2][m0 + gq + 8] = part1; }
    __syncthreads();                                // row sums complete; dV reads of P complete
    {
      const float D0 = s.red[0][m0 + gq] + s.red[1][m0 + gq], D1 = s.red[0][m0 + gq + 8] + s.red[1][m0 + gq + 8];
#pragma unroll
      for (int n = 0; n < NKB; ++n)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int i = m0 + gq + ((r & 2) ? 8 : 0), jl = n0 + nt * 8 + 2 * tq + (r & 1);
            float* pp = &s.p[i * S::PS + n * NTOK + jl];
            const float ds = *pp * (dp[n][nt][r] - ((r & 2) ? D1 : D0));
            *pp = ds;
            if (dtables) atomicAdd(&s.dbias[n][((i >> 3) - (jl >> 3) + 7) * 15 + ((i & 7) - (jl & 7) + 7)], ds);
          }
    }
    __syncthreads();
    // dQ = scale * dS . K ;  dK[kb] = scale * dS[:, kb]^T . Q
    attn::tile_pv<HD, false, NKB * NTOK, false>(s.p, S::PS, s.k, S::HS, dq, C, h * HD, s.rowq, scale, tid);
#pragma unroll
    for (int n = 0; n < NKB; ++n)
      attn::tile_pv<HD, true, NTOK, ATOMIC>(s.p + n * NTOK, S::PS, s.q, S::HS, dkv, 2 * C, h * HD, s.rowk[n], scale, tid);
  }
  __syncthreads();
  if (dtables) {
#pragma unroll
    for (int n = 0; n < NKB; ++n)
      for (int i = tid; i < 225; i += 256)
        atomicAdd(&dtables[((int64_t)(l1 * g.L + kb[n]) * 225 + i) * g.heads + h], s.dbias[n][i]);
  }
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
  FA_REQUIRE((int64_t)L * B * H * W < (1ll << 31) / 64, "%s: too many tokens", who);
  g.L = L; g.B = B; g.H = H; g.W = W; g.heads = heads; g.shift = shift; g.nWy = H / WIN; g.nWx = W / WIN; g.kind = kind;
  return FA_OK;
}

template <int HD, int NKB>
int launch_jfwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, float* o, const JGeom& g, float scale,
                const float* tables, cudaStream_t st) {
  const size_t smem = sizeof(JSmemF<HD, NKB>);
  FA_SMEM_ATTR_ONCE(smem, joint_fwd_kernel<HD, NKB>);
  dim3 grid((unsigned)(g.B * g.nWy * g.nWx * g.heads), (unsigned)g.L);
  joint_fwd_kernel<HD, NKB><<<grid, 256, smem, st>>>(q, ldq, kv, ldkv, o, g, scale, tables);
  FA_LAUNCH_CHECK("fa_joint_attn_fwd");
  return FA_OK;
}

template <int HD, int NKB, bool ATOMIC>
int launch_jbwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, const float* dout, float* dq, float* dkv,
                const JGeom& g, float scale, const float* tables, float* dtables, cudaStream_t st) {
  const size_t smem = sizeof(JSmemB<HD, NKB>);
  FA_SMEM_ATTR_ONCE(smem, joint_bwd_kernel<HD, NKB, ATOMIC>);
  const int items = g.B * g.nWy * g.nWx * g.heads;
  // persistent: exactly one resident wave (a 4-per-SM guess on a 3-per-SM kernel ran 1.33 waves = 2 rounds)
  static int per_sm = 0;
  if (!per_sm) {
    FA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, joint_bwd_kernel<HD, NKB, ATOMIC>, 256, smem));
    if (per_sm < 1) per_sm = 1;
  }
  int gx = (per_sm * kNumSMs / g.L / g.heads) * g.heads;                 // multiple of heads
  if (gx < g.heads) gx = g.heads;
  if (gx > items) gx = items;                                           // items is a multiple of heads
  dim3 grid((unsigned)gx, (unsigned)g.L);
  joint_bwd_kernel<HD, NKB, ATOMIC><<<grid, 256, smem, st>>>(q, ldq, kv, ldkv, dout, dq, dkv, g, scale, tables, dtables, items);
  FA_LAUNCH_CHECK("fa_joint_attn_bwd");
  return FA_OK;
}

}  // namespace

extern "C" {

int fa_joint_attn_fwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, float* o, int L, int B, int H, int W,
                      int heads, int hd, int shift, float scale, const float* tables, int kind, fa_stream_t stream) {
  FA_REQUIRE(q && kv && o && tables, "fa_joint_attn_fwd: null pointer");
  FA_REQUIRE(((uintptr_t)q | (uintptr_t)kv | (uintptr_t)o) % 16 == 0, "fa_joint_attn_fwd: pointers must be 16-byte aligned");
  JGeom g;
  int rc = jcheck("fa_joint_attn_fwd", L, B, H, W, heads, hd, shift, kind, ldq, ldkv, g);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_JOINT_ATTN, st);
  const int nkb = (kind == 0) ? 1 : L - 1;
  if (hd == 28) return nkb == 1 ? launch_jfwd<28, 1>(q, ldq, kv, ldkv, o, g, scale, tables, st)
                                : launch_jfwd<28, 2>(q, ldq, kv, ldkv, o, g, scale, tables, st);
  return nkb == 1 ? launch_jfwd<56, 1>(q, ldq, kv, ldkv, o, g, scale, tables, st)
                  : launch_jfwd<56, 2>(q, ldq, kv, ldkv, o, g, scale, tables, st);
}

int fa_joint_attn_bwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, const float* dout, float* dq,
                      float* dkv, int L, int B, int H, int W, int heads, int hd, int shift, float scale,
                      const float* tables, float* dtables, int kind, fa_stream_t stream) {
  FA_REQUIRE(q && kv && dout && dq && dkv && tables, "fa_joint_attn_bwd: null pointer");
  FA_REQUIRE(((uintptr_t)q | (uintptr_t)kv | (uintptr_t)dout | (uintptr_t)dq | (uintptr_t)dkv) % 16 == 0,
             "fa_joint_attn_bwd: pointers must be 16-byte aligned");
  JGeom g;
  int rc = jcheck("fa_joint_attn_bwd", L, B, H, W, heads, hd, shift, kind, ldq, ldkv, g);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_JOINT_ATTN, st);
  const int nkb = (kind == 0) ? 1 : L - 1;
  const int C = heads * hd;
  if (nkb == 2) {
    // two query bands add into every key band's gradient: zero-fill, then fp32 atomics
    FA_CUDA(cudaMemsetAsync(dkv, 0, (size_t)L * B * H * W * 2 * C * sizeof(float), st));
    if (hd == 28) return launch_jbwd<28, 2, true>(q, ldq, kv, ldkv, dout, dq, dkv, g, scale, tables, dtables, st);
    return launch_jbwd<56, 2, true>(q, ldq, kv, ldkv, dout, dq, dkv, g, scale, tables, dtables, st);
  }
  if (hd == 28) return launch_jbwd<28, 1, false>(q, ldq, kv, ldkv, dout, dq, dkv, g, scale, tables, dtables, st);
  return launch_jbwd<56, 1, false>(q, ldq, kv, ldkv, dout, dq, dkv, g, scale, tables, dtables, st);
}

}  // extern "C"
