// Training-pair generator on the device: the pixel work of TrainDataset.__getitem__
// (utils/dataset_utils.py:97-135) for a whole batch in one launch - synthetic Gaussian degradation in 0..255 space with
// clip + uint8 truncation (:122-126), the paired random crop (_crop_patch, :50-59), one of the 8 flip / rot90
// augmentations (image_utils.py:133-182) and ToTensor (HWC uint8 -> CHW float / 255).  The random DRAWS (image order,
// crop origin, augmentation id, sigma choice) stay on the host exactly as the reference makes them (a few integers per
// sample); the images live in HBM as one uint8 pool.  HBM-bound byte work: one thread per output pixel, 3 + 3 bytes in,
// 6 floats out, output writes coalesced along x.
#include "freqair_internal.h"

namespace {

// Counter-based noise: the reference adds ONE noise field to the whole image and cuts both patches of a pair from it,
// so overlapping pixels of the two crops share their noise.  A stateless generator keyed on (seed, absolute pixel,
// channel) reproduces that without materialising the field: two rounds of a 64-bit mix (splitmix64 finaliser) give
// 2 x 24 uniform bits, Box-Muller gives the normal.  oracle/datagen.py restates it bit for bit on the integer side.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ float normal_of(uint64_t seed, uint64_t counter) {
  const uint64_t h = mix64(mix64(seed) ^ counter);
  const float u1 = ((float)((h >> 40) & 0xFFFFFFu) + 1.0f) * (1.0f / 16777216.0f);      // (0, 1]
  const float u2 = (float)((h >> 8) & 0xFFFFFFu) * (1.0f / 16777216.0f);                // [0, 1)
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// meta[i] = {clean offset, degraded offset (-1: synthesise noise), H, W, y0, x0, mode, seed}
__global__ void __launch_bounds__(256) crop_augment_kernel(const uint8_t* __restrict__ pool, const int64_t* __restrict__ meta,
                                                           const float* __restrict__ sigma,
                                                           const float* __restrict__ noise, float* __restrict__ out_deg,
                                                           float* __restrict__ out_clean, int B, int P) {
  const int64_t total = (int64_t)B * P * P;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % P), y = (int)((i / P) % P), b = (int)(i / ((int64_t)P * P));
    const int64_t* m = meta + (int64_t)b * 8;
    const int W = (int)m[3], y0 = (int)m[4], x0 = (int)m[5], mode = (int)m[6];
    // out[y][x] = patch[py][px] for np.rot90 (counter-clockwise) k = mode/2 followed by flipud when mode is odd
    const int yy = (mode & 1) ? P - 1 - y : y;
    int py, px;
    switch (mode >> 1) {
      case 0: py = yy; px = x; break;
      case 1: py = x; px = P - 1 - yy; break;
      case 2: py = P - 1 - yy; px = P - 1 - x; break;
      default: py = P - 1 - x; px = yy; break;
    }
    const int sy = y0 + py, sx = x0 + px;
    const int64_t pix = (int64_t)sy * W + sx;
    const uint8_t* c = pool + m[0] + pix * 3;
    float cv[3], dv[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) cv[ch] = (float)c[ch];
    if (m[1] >= 0) {
      const uint8_t* d = pool + m[1] + pix * 3;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) dv[ch] = (float)d[ch];
    } else {
      const float sg = sigma[b];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float z = noise ? noise[(((int64_t)b * P + py) * P + px) * 3 + ch] : normal_of((uint64_t)m[7], (uint64_t)(pix * 3 + ch));
        const float v = fminf(fmaxf(cv[ch] + z * sg, 0.f), 255.f);
        dv[ch] = truncf(v);                                            // .astype(np.uint8)
      }
    }
    const int64_t plane = (int64_t)P * P;
    const int64_t o = (int64_t)b * 3 * plane + (int64_t)y * P + x;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      out_clean[o + ch * plane] = cv[ch] / 255.0f;                      // ToTensor
      out_deg[o + ch * plane] = dv[ch] / 255.0f;
    }
  }
}

}  // namespace

extern "C" {

int fa_crop_augment(const uint8_t* pool, const int64_t* meta, const float* sigma, const float* noise, float* out_degraded,
                    float* out_clean, int B, int P, fa_stream_t stream) {
  FA_REQUIRE(pool && meta && out_degraded && out_clean, "fa_crop_augment: null pointer");
  FA_REQUIRE(P >= 1 && B >= 0, "fa_crop_augment: bad shape");
  if (B == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  const int64_t total = (int64_t)B * P * P;
  int64_t grid = (total + 255) / 256;
  if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
  crop_augment_kernel<<<(unsigned)grid, 256, 0, st>>>(pool, meta, sigma, noise, out_degraded, out_clean, B, P);
  FA_LAUNCH_CHECK("fa_crop_augment");
  return FA_OK;
}

}  // extern "C"
