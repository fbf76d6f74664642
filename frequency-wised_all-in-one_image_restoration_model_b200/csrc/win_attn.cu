// K2 (decoder form) - LeWin window multi-head self-attention with the frequency-band re-weighting of the
// post-softmax attention map fused in.  One CTA (288 threads) per (image, window, head):
//   gather the window's 64 tokens straight from image-order q / kv (cyclic shift folded into the index
//   math, so roll + window_partition + window_reverse never touch HBM), S = scale*q.k^T + rel-pos bias
//   + shift mask, softmax, P' = irfft2(rfft2(P)*(1+coef[band])) in shared memory (fft64.cuh), O = P'.V,
//   scatter O back in image order.  The 64x64 map never leaves the SM; HBM traffic is q,k,v,o only.
// Backward recomputes P instead of saving 1.2 GB of maps: the filter is self-adjoint, d(coef) comes from
// Parseval on the two half-spectra, and the relative-position-bias gradient is reduced in shared memory
// across all windows a CTA visits before one flush of 225 atomics.
#include "freqair_internal.h"
#include "fft64.cuh"
#include "attn_tiles.cuh"

namespace {

constexpr int NTHR = 288;
using attn::WIN;
using attn::NTOK;
using attn::Pitch;

struct WinGeom {
  int B, H, W, heads, shift, nWy, nWx;
};

// token row (b*H*W + y*W + x) of window position p, and its SW-MSA region label
__device__ __forceinline__ void token_of(const WinGeom& g, int b, int wy, int wx, int p, int& row, int& label) {
  const int sy = wy * WIN + (p >> 3), sx = wx * WIN + (p & 7);
  int y = sy + g.shift, x = sx + g.shift;
  if (y >= g.H) y -= g.H;
  if (x >= g.W) x -= g.W;
  row = (b * g.H + y) * g.W + x;
  const int ry = sy < g.H - WIN ? 0 : (sy < g.H - g.shift ? 1 : 2);
  const int rx = sx < g.W - WIN ? 0 : (sx < g.W - g.shift ? 1 : 2);
  label = g.shift > 0 ? ry * 3 + rx : 0;
}

// Dropout on the attention map (nn.Dropout after the band re-weighting, encoder_ViT.py:94): keep-mask from a stateless
// hash of (seed, item, i, j) so that the backward kernel regenerates exactly the mask the forward applied.  The seed is
// read from DEVICE memory (one int64 drawn by torch's generator per call), so a captured CUDA graph draws a new mask at
// every replay.  Returns 0 or 1/keep.
__device__ __forceinline__ float drop_scale(uint64_t seed, uint32_t item, uint32_t ij, float p, float inv_keep) {
  uint64_t x = seed ^ (((uint64_t)item << 12) | ij) * 0x9E3779B97F4A7C15ull;       // splitmix64 finaliser
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  const float u = (float)(uint32_t)(x >> 40) * (1.0f / 16777216.0f);               // 24 uniform bits in [0, 1)
  return u < p ? 0.f : inv_keep;
}
__device__ __forceinline__ void drop_map(float* P, uint64_t seed, uint32_t item, float p, int tid) {
  const float inv_keep = 1.0f / (1.0f - p);
  for (int e = tid; e < NTOK * NTOK; e += NTHR) P[(e >> 6) * fft64::PSTR + (e & 63)] *= drop_scale(seed, item, e, p, inv_keep);
}

template <int HD>
struct Smem {
  static constexpr int HS = Pitch<HD>::HS;
  static constexpr int HSV = Pitch<HD>::HSV;
  float q[NTOK * HS];
  float k[NTOK * HS];
  float v[NTOK * HSV];
  float p[NTOK * fft64::PSTR];
  // the half spectrum of the filter phase reuses the q|k tiles (dead once the scores exist): 85 KB -> 68 KB, 3 CTAs per SM
  static_assert(2 * NTOK * HS * sizeof(float) >= NTOK * fft64::SPSTR * sizeof(float2), "q|k tiles too small for the spectrum");
  __device__ __forceinline__ float2* sp() { return reinterpret_cast<float2*>(q); }
  float bias[232];
  float coef[16];
  int row[NTOK];
  int label[NTOK];
  uint8_t band[NTOK * 33 + 8];
};

// thin adaptors onto the shared tensor-core tile routines (64 keys, map pitch fft64::PSTR)
template <int HD, int HS>
__device__ __forceinline__ void load_tile(float* dst, const float* __restrict__ src, int64_t ld, int col0,
                                          const int* rows, int tid) {
  attn::load_tile<HD, HS, NTHR>(dst, src, ld, col0, rows, tid);
}
template <int HD, bool BIASMASK>
__device__ __forceinline__ void tile_abt(const float* A, int lda, const float* Bm, int ldb, float* out, float scale,
                                         const float* bias, const int* label, int tid) {
  attn::tile_abt<HD, BIASMASK>(A, lda, Bm, ldb, out, fft64::PSTR, scale, bias, label, tid);
}
__device__ __forceinline__ void softmax_rows(float* P, int tid) { attn::softmax_rows<64>(P, fft64::PSTR, tid); }
template <int HD, bool TRANS>
__device__ __forceinline__ void tile_pv(const float* P, const float* V, int ldv, float* __restrict__ out, int64_t ld,
                                        int col0, const int* rows, float scale, int tid) {
  attn::tile_pv<HD, TRANS, 64, false>(P, fft64::PSTR, V, ldv, out, ld, col0, rows, scale, tid);
}

template <int HD>
__global__ void __launch_bounds__(NTHR) win_attn_fwd_kernel(const float* __restrict__ q, int64_t ldq,
                                                            const float* __restrict__ kv, int64_t ldkv,
                                                            float* __restrict__ o, WinGeom g, float scale,
                                                            const float* __restrict__ table,
                                                            const float* __restrict__ coef, int coef_bstride,
                                                            const uint8_t* __restrict__ band_of_bin, int nbands,
                                                            float drop_p, const int64_t* __restrict__ drop_seed) {
  extern __shared__ __align__(16) unsigned char smraw[];
  Smem<HD>& s = *reinterpret_cast<Smem<HD>*>(smraw);
  const int tid = threadIdx.x;
  const int C = g.heads * HD;
  int id = blockIdx.x;
  const int h = id % g.heads; id /= g.heads;
  const int wx = id % g.nWx; id /= g.nWx;
  const int wy = id % g.nWy;
  const int b = id / g.nWy;

  if (tid < NTOK) token_of(g, b, wy, wx, tid, s.row[tid], s.label[tid]);
  for (int i = tid; i < 225; i += NTHR) s.bias[i] = table ? table[i * g.heads + h] : 0.f;
  if (coef) {
    for (int i = tid; i < NTOK * 33; i += NTHR) s.band[i] = band_of_bin[i];
    if (tid < 16) s.coef[tid] = tid < nbands ? coef[((int64_t)b * coef_bstride + h) * nbands + tid] : 0.f;
  }
  __syncthreads();
  constexpr int HS = Smem<HD>::HS, HSV = Smem<HD>::HSV;
  load_tile<HD, HS>(s.q, q, ldq, h * HD, s.row, tid);
  load_tile<HD, HS>(s.k, kv, ldkv, h * HD, s.row, tid);
  load_tile<HD, HSV>(s.v, kv, ldkv, C + h * HD, s.row, tid);
  attn::load_wait();
  __syncthreads();
  tile_abt<HD, true>(s.q, HS, s.k, HS, s.p, scale, s.bias, s.label, tid);
  __syncthreads();
  softmax_rows(s.p, tid);
  __syncthreads();
  if (coef) fft64::filter_map(s.p, s.sp(), s.band, s.coef, 1.0f, tid);
  if (drop_seed) {
    drop_map(s.p, (uint64_t)__ldg(drop_seed), blockIdx.x, drop_p, tid);
    __syncthreads();
  }
  tile_pv<HD, false>(s.p, s.v, HSV, o, C, h * HD, s.row, 1.0f, tid);
}

// ------------------------------------------------------------------ backward
template <int HD>
struct SmemB {
  static constexpr int HS = Pitch<HD>::HS;
  // The two half-spectrum buffers of the filter phase live on top of the q|k|v tiles when those are large enough (hd 56,
  // 64): q, k, v are dead between the score / dP' contractions and the final dQ / dK ones, and q, k are simply gathered
  // again (L2 hits) afterwards.  That takes the CTA from 135 KB to 101 KB of shared memory, i.e. from 1 to 2 CTAs per SM
  // - the kernel is a chain of short phases separated by block barriers, so the second CTA roughly doubles throughput.
  static constexpr bool ALIAS = 3 * NTOK * HS * sizeof(float) >= 2 * NTOK * fft64::SPSTR * sizeof(float2);
  float q[NTOK * HS];
  float k[NTOK * HS];
  float v[NTOK * HS];
  float dO[NTOK * HS];
  float p[NTOK * fft64::PSTR];       // P (softmax), kept for dS
  float x[NTOK * fft64::PSTR];       // dP' -> P' -> dP -> dS
  float2 sp_own[ALIAS ? 1 : 2 * NTOK * fft64::SPSTR];
  __device__ __forceinline__ float2* spA() { return ALIAS ? reinterpret_cast<float2*>(q) : sp_own; }
  __device__ __forceinline__ float2* spB() { return spA() + NTOK * fft64::SPSTR; }
  float bias[232];
  float dbias[232];
  float coef[16];
  float ecoef[16];
  int row[NTOK];
  int label[NTOK];
  uint8_t band[NTOK * 33 + 8];
};

template <int HD>
__global__ void __launch_bounds__(NTHR) win_attn_bwd_kernel(const float* __restrict__ q, int64_t ldq,
                                                            const float* __restrict__ kv, int64_t ldkv,
                                                            const float* __restrict__ dout, float* __restrict__ dq,
                                                            float* __restrict__ dkv, WinGeom g, float scale,
                                                            const float* __restrict__ table, float* __restrict__ dtable,
                                                            const float* __restrict__ coef, int coef_bstride,
                                                            float* __restrict__ dcoef,
                                                            const uint8_t* __restrict__ band_of_bin, int nbands,
                                                            int total_items, float drop_p,
                                                            const int64_t* __restrict__ drop_seed) {
  extern __shared__ __align__(16) unsigned char smraw[];
  SmemB<HD>& s = *reinterpret_cast<SmemB<HD>*>(smraw);
  const int tid = threadIdx.x;
  const int C = g.heads * HD;
  // gridDim.x is a multiple of heads, so every item this CTA visits has the same head
  const int h = blockIdx.x % g.heads;
  for (int i = tid; i < 232; i += NTHR) { s.bias[i] = (table && i < 225) ? table[i * g.heads + h] : 0.f; s.dbias[i] = 0.f; }
  if (coef) for (int i = tid; i < NTOK * 33; i += NTHR) s.band[i] = band_of_bin[i];

  for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
    int id = item / g.heads;
    const int wx = id % g.nWx; id /= g.nWx;
    const int wy = id % g.nWy;
    const int b = id / g.nWy;
    __syncthreads();
    if (tid < NTOK) token_of(g, b, wy, wx, tid, s.row[tid], s.label[tid]);
    if (coef && tid < 16) {
      s.coef[tid] = tid < nbands ? coef[((int64_t)b * coef_bstride + h) * nbands + tid] : 0.f;
      s.ecoef[tid] = 0.f;
    }
    __syncthreads();
    constexpr int HS = SmemB<HD>::HS;
    load_tile<HD, HS>(s.q, q, ldq, h * HD, s.row, tid);
    load_tile<HD, HS>(s.k, kv, ldkv, h * HD, s.row, tid);
    load_tile<HD, HS>(s.v, kv, ldkv, C + h * HD, s.row, tid);
    load_tile<HD, HS>(s.dO, dout, C, h * HD, s.row, tid);
    attn::load_wait();
    __syncthreads();
    tile_abt<HD, true>(s.q, HS, s.k, HS, s.p, scale, s.bias, s.label, tid);     // S
    tile_abt<HD, false>(s.dO, HS, s.v, HS, s.x, 1.0f, nullptr, nullptr, tid);   // dP' = dO.V^T
    __syncthreads();
    softmax_rows(s.p, tid);                                             // P
    if (drop_seed) drop_map(s.x, (uint64_t)__ldg(drop_seed), item, drop_p, tid);    // dropout backward: dP' o mask / keep
    __syncthreads();
    if (coef) {
      // F(dP') -> spB
      fft64::rows_forward(s.x, s.spB(), tid);
      __syncthreads();
      {
        const int col = min(tid >> 3, 32), l = tid & 7;
        float2 a[8];
        fft64::col_load(s.spB(), a, col, l);
        fft64::fft64_group<-1>(a, l);
        __syncwarp();
        if ((tid >> 3) <= 32) fft64::col_store(s.spB(), a, col, l);
      }
      fft64::rows_forward(s.p, s.spA(), tid);
      __syncthreads();
      {
        // F(P): band energies against F(dP'), then gain and inverse columns -> spA
        const int col = min(tid >> 3, 32), l = tid & 7;
        const bool live = (tid >> 3) <= 32;
        float2 a[8], bb[8];
        fft64::col_load(s.spA(), a, col, l);
        fft64::col_load(s.spB(), bb, col, l);
        fft64::fft64_group<-1>(a, l);
        const float wgt = (col == 0 || col == 32) ? 1.0f : 2.0f;
        float e[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) e[t] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int bnd = s.band[(l + 8 * j) * 33 + col];
          const float dot = wgt * (a[j].x * bb[j].x + a[j].y * bb[j].y);
#pragma unroll
          for (int t = 0; t < 8; ++t) e[t] += (bnd == t) ? dot : 0.f;
          const float gn = 1.0f + s.coef[bnd];
          a[j].x *= gn; a[j].y *= gn;
        }
        fft64::fft64_group<1>(a, l);
        __syncwarp();
        if (live) fft64::col_store(s.spA(), a, col, l);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          float v = live ? e[t] : 0.f;
          v = warp_sum(v);
          if ((tid & 31) == 0 && t < nbands) atomicAdd(&s.ecoef[t], v * (1.0f / 4096.0f));
        }
      }
      __syncthreads();
      fft64::rows_inverse(s.spA(), s.x, 1.0f / 4096.0f, tid);             // x = P'
      __syncthreads();
      if (drop_seed) {                                                  // the forward multiplied P' by the same mask
        drop_map(s.x, (uint64_t)__ldg(drop_seed), item, drop_p, tid);
        __syncthreads();
      }
      tile_pv<HD, true>(s.x, s.dO, HS, dkv, 2 * C, C + h * HD, s.row, 1.0f, tid);   // dV = P'^T.dO
      {
        // dP = filter(dP'): gain on F(dP'), inverse columns
        const int col = min(tid >> 3, 32), l = tid & 7;
        float2 a[8];
        fft64::col_load(s.spB(), a, col, l);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gn = 1.0f + s.coef[s.band[(l + 8 * j) * 33 + col]];
          a[j].x *= gn; a[j].y *= gn;
        }
        fft64::fft64_group<1>(a, l);
        __syncwarp();
        if ((tid >> 3) <= 32) fft64::col_store(s.spB(), a, col, l);
      }
      __syncthreads();                                                  // dV reads of x done, spB complete
      fft64::rows_inverse(s.spB(), s.x, 1.0f / 4096.0f, tid);             // x = dP
      __syncthreads();
      if (SmemB<HD>::ALIAS) {            // the spectra overwrote q|k|v: gather q and k again for dQ / dK
        load_tile<HD, HS>(s.q, q, ldq, h * HD, s.row, tid);
        load_tile<HD, HS>(s.k, kv, ldkv, h * HD, s.row, tid);
      }
      if (dcoef && tid < nbands) atomicAdd(&dcoef[((int64_t)b * coef_bstride + h) * nbands + tid], s.ecoef[tid]);
    } else {
      tile_pv<HD, true>(s.p, s.dO, HS, dkv, 2 * C, C + h * HD, s.row, 1.0f, tid);   // dV = P^T.dO
      __syncthreads();
    }
    // dS = P o (dP - rowsum(dP o P)) -> x ; bias gradient into shared memory
    if (tid < 256) {
      const int w = tid >> 5, lane = tid & 31;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int i = w * 8 + r;
        float* xr = s.x + i * fft64::PSTR;
        const float* pr = s.p + i * fft64::PSTR;
        const float p0 = pr[lane], p1 = pr[lane + 32];
        const float d0 = xr[lane], d1 = xr[lane + 32];
        const float dotv = warp_sum(p0 * d0 + p1 * d1);
        const float s0 = p0 * (d0 - dotv), s1 = p1 * (d1 - dotv);
        xr[lane] = s0; xr[lane + 32] = s1;
        if (dtable) {
          const int j0 = lane, j1 = lane + 32;
          atomicAdd(&s.dbias[((i >> 3) - (j0 >> 3) + 7) * 15 + ((i & 7) - (j0 & 7) + 7)], s0);
          atomicAdd(&s.dbias[((i >> 3) - (j1 >> 3) + 7) * 15 + ((i & 7) - (j1 & 7) + 7)], s1);
        }
      }
    }
    attn::load_wait();                                                  // the q|k reload (ALIAS) overlapped the dS pass
    __syncthreads();
    tile_pv<HD, false>(s.x, s.k, HS, dq, C, h * HD, s.row, scale, tid);        // dQ = scale * dS.K
    tile_pv<HD, true>(s.x, s.q, HS, dkv, 2 * C, h * HD, s.row, scale, tid);    // dK = scale * dS^T.Q
  }
  __syncthreads();
  if (dtable) for (int i = tid; i < 225; i += NTHR) atomicAdd(&dtable[i * g.heads + h], s.dbias[i]);
}

template <int HD>
int launch_fwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, float* o, const WinGeom& g, float scale,
               const float* table, const float* coef, int cbs, const uint8_t* bob, int nbands, float drop_p,
               const int64_t* drop_seed, cudaStream_t st) {
  const size_t smem = sizeof(Smem<HD>);
  FA_CUDA(cudaFuncSetAttribute(win_attn_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int items = g.B * g.nWy * g.nWx * g.heads;
  win_attn_fwd_kernel<HD><<<items, NTHR, smem, st>>>(q, ldq, kv, ldkv, o, g, scale, table, coef, cbs, bob, nbands, drop_p, drop_seed);
  FA_LAUNCH_CHECK("fa_win_attn_fwd");
  return FA_OK;
}

template <int HD>
int launch_bwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, const float* dout, float* dq, float* dkv,
               const WinGeom& g, float scale, const float* table, float* dtable, const float* coef, int cbs,
               float* dcoef, const uint8_t* bob, int nbands, float drop_p, const int64_t* drop_seed, cudaStream_t st) {
  const size_t smem = sizeof(SmemB<HD>);
  FA_CUDA(cudaFuncSetAttribute(win_attn_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int items = g.B * g.nWy * g.nWx * g.heads;
  static int per_sm = 0;                       // persistent: exactly one resident wave
  if (!per_sm) {
    FA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, win_attn_bwd_kernel<HD>, NTHR, smem));
    if (per_sm < 1) per_sm = 1;
  }
  int grid = (per_sm * kNumSMs / g.heads) * g.heads;
  if (grid < g.heads) grid = g.heads;
  if (grid > items) grid = items;          // items is a multiple of heads
  win_attn_bwd_kernel<HD><<<grid, NTHR, smem, st>>>(q, ldq, kv, ldkv, dout, dq, dkv, g, scale, table, dtable, coef, cbs,
                                                   dcoef, bob, nbands, items, drop_p, drop_seed);
  FA_LAUNCH_CHECK("fa_win_attn_bwd");
  return FA_OK;
}

int check_geom(const char* who, int B, int H, int W, int heads, int hd, int shift, int64_t ldq, int64_t ldkv,
               const void* coef, const void* bob, int nbands, WinGeom& g) {
  FA_REQUIRE(B > 0 && heads > 0, "%s: empty batch/heads", who);
  FA_REQUIRE(H % WIN == 0 && W % WIN == 0, "%s: H=%d W=%d must be multiples of the 8x8 window", who, H, W);
  FA_REQUIRE(hd == 28 || hd == 56 || hd == 64, "%s: head_dim=%d unsupported (28, 56, 64)", who, hd);
  FA_REQUIRE(shift == 0 || shift == 4, "%s: shift=%d unsupported (0 or 4)", who, shift);
  FA_REQUIRE(shift == 0 || (H > WIN && W > WIN), "%s: shifted windows need H,W > 8", who);
  FA_REQUIRE(ldq % 4 == 0 && ldkv % 4 == 0, "%s: row strides must be multiples of 4 floats", who);
  FA_REQUIRE((int64_t)B * H * W < (1ll << 31) / 64, "%s: too many tokens", who);
  FA_REQUIRE(!coef || (bob && nbands >= 1 && nbands <= 8), "%s: coef needs band_of_bin and 1..8 bands", who);
  g.B = B; g.H = H; g.W = W; g.heads = heads; g.shift = shift; g.nWy = H / WIN; g.nWx = W / WIN;
  return FA_OK;
}

}  // namespace

extern "C" {

int fa_win_attn_fwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, float* o, int B, int H, int W, int heads,
                    int hd, int shift, float scale, const float* table, const float* coef, int coef_bstride,
                    const uint8_t* band_of_bin, int nbands, float drop_p, const int64_t* drop_seed, fa_stream_t stream) {
  FA_REQUIRE(q && kv && o, "fa_win_attn_fwd: null pointer");
  FA_REQUIRE(!drop_seed || (drop_p > 0.f && drop_p < 1.f && coef), "fa_win_attn_fwd: dropout needs 0 < p < 1 and the coef path");
  FA_REQUIRE(((uintptr_t)q | (uintptr_t)kv | (uintptr_t)o) % 16 == 0, "fa_win_attn_fwd: pointers must be 16-byte aligned");
  WinGeom g;
  int rc = check_geom("fa_win_attn_fwd", B, H, W, heads, hd, shift, ldq, ldkv, coef, band_of_bin, nbands, g);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_WIN_ATTN, st);
  if (hd == 56) return launch_fwd<56>(q, ldq, kv, ldkv, o, g, scale, table, coef, coef_bstride, band_of_bin, nbands, drop_p, drop_seed, st);
  if (hd == 28) return launch_fwd<28>(q, ldq, kv, ldkv, o, g, scale, table, coef, coef_bstride, band_of_bin, nbands, drop_p, drop_seed, st);
  return launch_fwd<64>(q, ldq, kv, ldkv, o, g, scale, table, coef, coef_bstride, band_of_bin, nbands, drop_p, drop_seed, st);
}

int fa_win_attn_bwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, const float* dout, float* dq, float* dkv,
                    int B, int H, int W, int heads, int hd, int shift, float scale, const float* table, float* dtable,
                    const float* coef, int coef_bstride, float* dcoef, const uint8_t* band_of_bin, int nbands,
                    float drop_p, const int64_t* drop_seed, fa_stream_t stream) {
  FA_REQUIRE(q && kv && dout && dq && dkv, "fa_win_attn_bwd: null pointer");
  FA_REQUIRE(!drop_seed || (drop_p > 0.f && drop_p < 1.f && coef), "fa_win_attn_bwd: dropout needs 0 < p < 1 and the coef path");
  FA_REQUIRE(((uintptr_t)q | (uintptr_t)kv | (uintptr_t)dout) % 16 == 0, "fa_win_attn_bwd: pointers must be 16-byte aligned");
  WinGeom g;
  int rc = check_geom("fa_win_attn_bwd", B, H, W, heads, hd, shift, ldq, ldkv, coef, band_of_bin, nbands, g);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_WIN_ATTN, st);
  if (hd == 56) return launch_bwd<56>(q, ldq, kv, ldkv, dout, dq, dkv, g, scale, table, dtable, coef, coef_bstride, dcoef, band_of_bin, nbands, drop_p, drop_seed, st);
  if (hd == 28) return launch_bwd<28>(q, ldq, kv, ldkv, dout, dq, dkv, g, scale, table, dtable, coef, coef_bstride, dcoef, band_of_bin, nbands, drop_p, drop_seed, st);
  return launch_bwd<64>(q, ldq, kv, ldkv, dout, dq, dkv, g, scale, table, dtable, coef, coef_bstride, dcoef, band_of_bin, nbands, drop_p, drop_seed, st);
}

}  // extern "C"
