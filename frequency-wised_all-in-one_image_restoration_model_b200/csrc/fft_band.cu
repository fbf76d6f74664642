// K1 - batched 2-D real FFT band filter / band split on n x n maps staged in shared memory.
// n = 64 (attention maps, the hot size) uses the radix-8 register/shuffle transforms of fft64.cuh;
// every other power of two up to 128 (128 = input crops, small sizes for tests) uses a compact radix-2
// shared-memory transform, one warp per 1-D FFT.  One CTA per map: the map is read from HBM once and
// written once per output band (algorithmic bytes 2*4*n*n per map per pass).
#include "freqair_internal.h"
#include "fft64.cuh"

namespace {

using fft64::cmul;

// ------------------------------------------------------------------ generic radix-2 path
struct GenSmem {
  float* re;        // [n][n+1]
  float2* sp;       // [n][nh]   nh = n/2+1
  float2* zw;       // [warps][n] per-warp scratch
  float2* tw;       // [n/2] forward twiddles exp(-2*pi*i*k/n)
};

__device__ __forceinline__ int brev(int i, int logn) { return (int)(__brev((unsigned)i) >> (32 - logn)); }

// in-place radix-2 DIT on d[0..n) with element stride `stride`; one warp; SIGN=-1 forward, +1 inverse
template <int SIGN>
__device__ void fft_warp(float2* d, int stride, int n, int logn, const float2* tw, int lane) {
  for (int i = lane; i < n; i += 32) {
    int j = brev(i, logn);
    if (i < j) { float2 t = d[i * stride]; d[i * stride] = d[j * stride]; d[j * stride] = t; }
  }
  __syncwarp();
  for (int s = 1; s <= logn; ++s) {
    const int m = 1 << s, half = m >> 1, tstep = n >> s;
    for (int idx = lane; idx < (n >> 1); idx += 32) {
      const int grp = idx / half, pos = idx % half;
      const int i0 = grp * m + pos, i1 = i0 + half;
      float2 w = tw[pos * tstep];
      if (SIGN > 0) w.y = -w.y;
      const float2 t = cmul(w, d[i1 * stride]);
      const float2 u = d[i0 * stride];
      d[i0 * stride] = make_float2(u.x + t.x, u.y + t.y);
      d[i1 * stride] = make_float2(u.x - t.x, u.y - t.y);
    }
    __syncwarp();
  }
}

__device__ void gen_forward(const GenSmem& s, int n, int logn) {
  const int nh = n / 2 + 1, rs = n + 1;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float2* z = s.zw + w * n;
  for (int r = w; r < n / 2; r += nw) {
    for (int i = lane; i < n; i += 32) z[i] = make_float2(s.re[(2 * r) * rs + i], s.re[(2 * r + 1) * rs + i]);
    __syncwarp();
    fft_warp<-1>(z, 1, n, logn, s.tw, lane);
    for (int k = lane; k < nh; k += 32) {
      const float2 a = z[k], b = z[(n - k) & (n - 1)];
      s.sp[(2 * r) * nh + k] = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
      s.sp[(2 * r + 1) * nh + k] = make_float2(0.5f * (a.y + b.y), 0.5f * (b.x - a.x));
    }
    __syncwarp();
  }
  __syncthreads();
  for (int c = w; c < nh; c += nw) fft_warp<-1>(s.sp + c, nh, n, logn, s.tw, lane);
  __syncthreads();
}

__device__ void gen_inverse(const GenSmem& s, int n, int logn) {
  const int nh = n / 2 + 1, rs = n + 1;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = w; c < nh; c += nw) fft_warp<1>(s.sp + c, nh, n, logn, s.tw, lane);
  __syncthreads();
  float2* z = s.zw + w * n;
  const float norm = 1.0f / ((float)n * (float)n);
  for (int r = w; r < n / 2; r += nw) {
    for (int k = lane; k < n; k += 32) {
      float2 xa, xb;
      if (k < nh) { xa = s.sp[(2 * r) * nh + k]; xb = s.sp[(2 * r + 1) * nh + k]; }
      else { xa = s.sp[(2 * r) * nh + n - k]; xb = s.sp[(2 * r + 1) * nh + n - k]; xa.y = -xa.y; xb.y = -xb.y; }
      z[k] = make_float2(xa.x - xb.y, xa.y + xb.x);
    }
    __syncwarp();
    fft_warp<1>(z, 1, n, logn, s.tw, lane);
    for (int i = lane; i < n; i += 32) {
      s.re[(2 * r) * rs + i] = z[i].x * norm;
      s.re[(2 * r + 1) * rs + i] = z[i].y * norm;
    }
    __syncwarp();
  }
  __syncthreads();
}

__device__ GenSmem gen_carve(float* base, int n) {
  GenSmem s;
  const int nh = n / 2 + 1;
  s.re = base;
  float* p = base + n * (n + 1);
  p += ((uintptr_t)p & 7) ? 1 : 0;
  s.sp = reinterpret_cast<float2*>(p);
  s.zw = s.sp + n * nh;
  s.tw = s.zw + (blockDim.x >> 5) * n;
  for (int k = threadIdx.x; k < n / 2; k += blockDim.x) {
    float sn, cs;
    sincospif(-2.0f * (float)k / (float)n, &sn, &cs);
    s.tw[k] = make_float2(cs, sn);
  }
  return s;
}

size_t gen_smem_bytes(int n, int threads) {
  const int nh = n / 2 + 1;
  return sizeof(float) * ((size_t)n * (n + 1) + 2) + sizeof(float2) * ((size_t)n * nh + (threads / 32) * n + n / 2);
}

// mode 0: y[band][map] spatial; mode 1: y[band][map][n][n][2] full spectrum; mode 2: filter with coef
__global__ void __launch_bounds__(256) band_generic_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           int64_t nmaps, int n, int logn,
                                                           const uint8_t* __restrict__ band_of_bin, int nbands, int mode,
                                                           const float* __restrict__ coef, int maps_per_group,
                                                           int heads) {
  extern __shared__ __align__(16) float smem[];
  GenSmem s = gen_carve(smem, n);
  const int nh = n / 2 + 1, rs = n + 1;
  const int64_t map = blockIdx.x;
  const int band = blockIdx.y;
  const float* xm = x + map * n * n;
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) s.re[(i / n) * rs + (i % n)] = xm[i];
  __syncthreads();
  gen_forward(s, n, logn);
  const float* cf = nullptr;
  if (mode == 2) cf = coef + ((map / maps_per_group) * heads + map % heads) * nbands;
  for (int i = threadIdx.x; i < n * nh; i += blockDim.x) {
    const int bnd = band_of_bin[i];
    const float g = (mode == 2) ? 1.0f + cf[bnd] : (bnd == band ? 1.0f : 0.0f);
    s.sp[i].x *= g; s.sp[i].y *= g;
  }
  __syncthreads();
  if (mode == 1) {
    float* ym = y + ((int64_t)band * nmaps + map) * n * n * 2;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
      const int u = i / n, k = i % n;
      float2 v;
      if (k < nh) v = s.sp[u * nh + k];
      else { v = s.sp[((n - u) & (n - 1)) * nh + (n - k)]; v.y = -v.y; }
      ym[2 * i] = v.x; ym[2 * i + 1] = v.y;
    }
    return;
  }
  gen_inverse(s, n, logn);
  float* ym = y + ((mode == 2) ? map : ((int64_t)band * nmaps + map)) * n * n;
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) ym[i] = s.re[(i / n) * rs + (i % n)];
}

// <a, band_i(b)> accumulated per (group, band):  (1/n^2) sum_bins wgt * Re(conj(A) B)
__global__ void __launch_bounds__(256) band_energy_generic_kernel(const float* __restrict__ a,
                                                                  const float* __restrict__ b, float* __restrict__ out,
                                                                  int n, int logn,
                                                                  const uint8_t* __restrict__ band_of_bin, int nbands,
                                                                  int maps_per_group, int heads) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float ebin[16];
  GenSmem s = gen_carve(smem, n);
  const int nh = n / 2 + 1, rs = n + 1;
  float2* spA = s.sp + (size_t)n * nh + (blockDim.x >> 5) * n + n / 2;      // second spectrum after the first carve
  const int64_t map = blockIdx.x;
  if (threadIdx.x < 16) ebin[threadIdx.x] = 0.f;
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) s.re[(i / n) * rs + (i % n)] = a[map * n * n + i];
  __syncthreads();
  gen_forward(s, n, logn);
  for (int i = threadIdx.x; i < n * nh; i += blockDim.x) spA[i] = s.sp[i];
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) s.re[(i / n) * rs + (i % n)] = b[map * n * n + i];
  __syncthreads();
  gen_forward(s, n, logn);
  const float norm = 1.0f / ((float)n * (float)n);
  for (int i = threadIdx.x; i < n * nh; i += blockDim.x) {
    const int k = i % nh;
    const float wgt = (k == 0 || k == n / 2) ? 1.0f : 2.0f;
    const float2 A = spA[i], Bv = s.sp[i];
    atomicAdd(&ebin[band_of_bin[i]], wgt * norm * (A.x * Bv.x + A.y * Bv.y));
  }
  __syncthreads();
  if (threadIdx.x < nbands)
    atomicAdd(&out[((map / maps_per_group) * heads + map % heads) * nbands + threadIdx.x], ebin[threadIdx.x]);
}


// Spectral L1 (train.py:70,91): mean | decompose(a) - decompose(b) | with inverse=False.  The band masks partition the
// spectrum and the FFT is linear, so the [nb, maps, n, n, 2] stack never exists: with F = fft2(a - b)
//   loss = (1 / (nb * maps * n^2 * 2)) * sum over banded bins (|Re F| + |Im F|)
//   dloss/da = Re(sum_k (sign Re F_k + i sign Im F_k) e^{+i theta}) = n^2 * irfft2(sign spectrum)   (Hermitian)
// One CTA per map: read a, b once, write the gradient once.
__global__ void __launch_bounds__(256) spectral_l1_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          float* __restrict__ loss, float* __restrict__ grad,
                                                          int64_t nmaps, int n, int logn,
                                                          const uint8_t* __restrict__ band_of_bin, int nbands,
                                                          float gscale) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[8];
  GenSmem s = gen_carve(smem, n);
  const int nh = n / 2 + 1, rs = n + 1;
  const int64_t map = blockIdx.x;
  const float* am = a + map * n * n;
  const float* bm = b + map * n * n;
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) s.re[(i / n) * rs + (i % n)] = am[i] - bm[i];
  __syncthreads();
  gen_forward(s, n, logn);
  float acc = 0.f;
  for (int i = threadIdx.x; i < n * nh; i += blockDim.x) {
    const int u = i / nh, k = i % nh;
    const bool edge = (k == 0 || k == n / 2);
    float2 v = s.sp[i];
    if (edge && (u == 0 || u == n / 2)) v.y = 0.f;          // self-conjugate bins are real
    if (band_of_bin[i] >= nbands) v = make_float2(0.f, 0.f);
    acc += (edge ? 1.0f : 2.0f) * (fabsf(v.x) + fabsf(v.y));
    s.sp[i] = make_float2(v.x > 0.f ? 1.f : (v.x < 0.f ? -1.f : 0.f), v.y > 0.f ? 1.f : (v.y < 0.f ? -1.f : 0.f));
  }
  const float inv = 1.0f / ((float)nbands * (float)nmaps * (float)n * (float)n * 2.0f);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    atomicAdd(loss, t * inv);
  }
  if (!grad) return;
  gen_inverse(s, n, logn);                                   // normalised by 1/n^2
  const float gs = gscale * inv * (float)n * (float)n;
  float* gm = grad + map * n * n;
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) gm[i] = s.re[(i / n) * rs + (i % n)] * gs;
}

// ------------------------------------------------------------------ n = 64 fast path (288 threads)
__global__ void __launch_bounds__(288) band64_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t nmaps,
                                                     const uint8_t* __restrict__ band_of_bin, int nbands, int mode,
                                                     const float* __restrict__ coef, int maps_per_group, int heads) {
  __shared__ float P[64 * fft64::PSTR];
  __shared__ float2 sp[64 * fft64::SPSTR];
  __shared__ uint8_t bandS[64 * 33];
  __shared__ float cf[16];
  const int tid = threadIdx.x;
  const int64_t map = blockIdx.x;
  const int band = blockIdx.y;
  const float* xm = x + map * 4096;
  for (int i = tid; i < 4096; i += 288) P[(i >> 6) * fft64::PSTR + (i & 63)] = xm[i];
  for (int i = tid; i < 64 * 33; i += 288) bandS[i] = band_of_bin[i];
  if (tid < 16) {
    float c = 0.f;
    if (tid < nbands) {
      if (mode == 2) c = coef[((map / maps_per_group) * heads + map % heads) * nbands + tid];
      else c = (tid == band) ? 1.0f : 0.0f;
    }
    cf[tid] = c;
  }
  __syncthreads();
  fft64::filter_map(P, sp, bandS, cf, mode == 2 ? 1.0f : 0.0f, tid);
  float* ym = y + ((mode == 2) ? map : ((int64_t)band * nmaps + map)) * 4096;
  for (int i = tid; i < 4096; i += 288) ym[i] = P[(i >> 6) * fft64::PSTR + (i & 63)];
}

__global__ void __launch_bounds__(256) dc_split_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t nmaps,
                                                       int nn) {
  __shared__ float sh[8];
  const int64_t map = blockIdx.x;
  const float* xm = x + map * nn;
  float s = 0.f;
  for (int i = threadIdx.x; i < nn; i += 256) s += xm[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += sh[i];
  const float mean = t / (float)nn;
  float* y0 = y + map * nn;
  float* y1 = y + (nmaps + map) * nn;
  for (int i = threadIdx.x; i < nn; i += 256) { y0[i] = mean; y1[i] = xm[i] - mean; }
}

int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }
bool pow2_ok(int n) { return n >= 8 && n <= 128 && (n & (n - 1)) == 0; }

int launch_generic(const float* x, float* y, int64_t nmaps, int n, const uint8_t* bob, int nbands, int mode, int nby,
                   const float* coef, int mpg, int heads, cudaStream_t st) {
  const size_t smem = gen_smem_bytes(n, 256);
  FA_CUDA(cudaFuncSetAttribute(band_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  band_generic_kernel<<<dim3((unsigned)nmaps, nby), 256, smem, st>>>(x, y, nmaps, n, ilog2(n), bob, nbands, mode, coef,
                                                                     mpg, heads);
  FA_LAUNCH_CHECK("fa_band(generic)");
  return FA_OK;
}

}  // namespace

extern "C" {

int fa_band_split(const float* x, float* y, int64_t nmaps, int n, const uint8_t* band_of_bin, int nbands, int mode,
                  fa_stream_t stream) {
  FA_REQUIRE(x && y && band_of_bin, "fa_band_split: null pointer");
  FA_REQUIRE(pow2_ok(n), "fa_band_split: n=%d unsupported (power of two in 8..128)", n);
  FA_REQUIRE(nbands >= 1 && nbands <= 16, "fa_band_split: nbands=%d unsupported (1..16)", nbands);
  FA_REQUIRE(mode == 0 || mode == 1, "fa_band_split: mode must be 0 (spatial) or 1 (spectrum)");
  FA_REQUIRE(nmaps < (1ll << 31), "fa_band_split: too many maps");
  if (nmaps == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_BAND_FILTER, st);
  if (n == 64 && mode == 0) {
    band64_kernel<<<dim3((unsigned)nmaps, nbands), 288, 0, st>>>(x, y, nmaps, band_of_bin, nbands, 0, nullptr, 1, 1);
    FA_LAUNCH_CHECK("fa_band_split(64)");
    return FA_OK;
  }
  return launch_generic(x, y, nmaps, n, band_of_bin, nbands, mode, nbands, nullptr, 1, 1, st);
}

int fa_band_filter(const float* x, float* y, int64_t nmaps, int n, const uint8_t* band_of_bin, int nbands,
                   const float* coef, int maps_per_group, int heads, fa_stream_t stream) {
  FA_REQUIRE(x && y && band_of_bin && coef, "fa_band_filter: null pointer");
  FA_REQUIRE(pow2_ok(n), "fa_band_filter: n=%d unsupported (power of two in 8..128)", n);
  FA_REQUIRE(nbands >= 1 && nbands <= 16, "fa_band_filter: nbands=%d unsupported (1..16)", nbands);
  FA_REQUIRE(maps_per_group >= 1 && heads >= 1, "fa_band_filter: bad grouping");
  FA_REQUIRE(nmaps < (1ll << 31), "fa_band_filter: too many maps");
  if (nmaps == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_BAND_FILTER, st);
  if (n == 64) {
    band64_kernel<<<dim3((unsigned)nmaps, 1), 288, 0, st>>>(x, y, nmaps, band_of_bin, nbands, 2, coef, maps_per_group, heads);
    FA_LAUNCH_CHECK("fa_band_filter(64)");
    return FA_OK;
  }
  return launch_generic(x, y, nmaps, n, band_of_bin, nbands, 2, 1, coef, maps_per_group, heads, st);
}

int fa_band_energy(const float* a, const float* b, float* out, int64_t nmaps, int n, const uint8_t* band_of_bin,
                   int nbands, int maps_per_group, int heads, fa_stream_t stream) {
  FA_REQUIRE(a && b && out && band_of_bin, "fa_band_energy: null pointer");
  FA_REQUIRE(pow2_ok(n), "fa_band_energy: n=%d unsupported (power of two in 8..128)", n);
  FA_REQUIRE(nbands >= 1 && nbands <= 16, "fa_band_energy: nbands=%d unsupported (1..16)", nbands);
  FA_REQUIRE(nmaps < (1ll << 31), "fa_band_energy: too many maps");
  if (nmaps == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_BAND_FILTER, st);
  const size_t smem = gen_smem_bytes(n, 256) + sizeof(float2) * (size_t)n * (n / 2 + 1);
  FA_CUDA(cudaFuncSetAttribute(band_energy_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  band_energy_generic_kernel<<<(unsigned)nmaps, 256, smem, st>>>(a, b, out, n, ilog2(n), band_of_bin, nbands,
                                                                 maps_per_group, heads);
  FA_LAUNCH_CHECK("fa_band_energy");
  return FA_OK;
}

int fa_spectral_l1(const float* a, const float* b, float* loss, float* grad, int64_t nmaps, int n,
                   const uint8_t* band_of_bin, int nbands, float gscale, fa_stream_t stream) {
  FA_REQUIRE(a && b && loss && band_of_bin, "fa_spectral_l1: null pointer");
  FA_REQUIRE(pow2_ok(n), "fa_spectral_l1: n=%d unsupported (power of two in 8..128)", n);
  FA_REQUIRE(nbands >= 1 && nbands <= 255, "fa_spectral_l1: nbands=%d unsupported (1..255)", nbands);
  FA_REQUIRE(nmaps > 0 && nmaps < (1ll << 31), "fa_spectral_l1: bad map count");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_BAND_FILTER, st);
  const size_t smem = gen_smem_bytes(n, 256);
  FA_CUDA(cudaFuncSetAttribute(spectral_l1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  spectral_l1_kernel<<<(unsigned)nmaps, 256, smem, st>>>(a, b, loss, grad, nmaps, n, ilog2(n), band_of_bin, nbands, gscale);
  FA_LAUNCH_CHECK("fa_spectral_l1");
  return FA_OK;
}

int fa_dc_split(const float* x, float* y, int64_t nmaps, int n, fa_stream_t stream) {
  FA_REQUIRE(x && y && n > 0, "fa_dc_split: bad argument");
  if (nmaps == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_BAND_FILTER, st);
  dc_split_kernel<<<(unsigned)nmaps, 256, 0, st>>>(x, y, nmaps, n * n);
  FA_LAUNCH_CHECK("fa_dc_split");
  return FA_OK;
}

}  // extern "C"
