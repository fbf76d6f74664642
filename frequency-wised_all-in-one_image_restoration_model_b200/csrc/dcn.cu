// K7 - DCNv2 (modulated deformable conv 3x3, stride 1, pad 1, one deformable group) on NHWC tokens.
// The bilinear gather is written as an im2col whose output feeds the dense contraction (fa_gemm):
// a warp owns one (pixel, tap) and sweeps the C channels with 128-bit loads, so every corner read is a
// contiguous, coalesced C*4-byte segment (NHWC is what makes the gather bandwidth-friendly).
// ref: net/utils/deform_conv.py:56-67 (+ the absent mmcv modulated_deform_conv2d; parity unpinned).
#include "freqair_internal.h"

namespace {

struct Tap { int y0, x0; float wy1, wx1; float w00, w01, w10, w11; bool v00, v01, v10, v11; float m; };

// om row layout (27 floats): raw conv_offset_mask output. The reference builds offset = cat(o1, o2)
// with o1 = ch 0..8, o2 = ch 9..17, so offset channel j = om[j]; tap k uses (dy, dx) = offset[2k], offset[2k+1].
__device__ __forceinline__ Tap make_tap(const float* omr, int k, int y, int x, int H, int W) {
  Tap t;
  const float dy = omr[2 * k], dx = omr[2 * k + 1];
  t.m = 1.0f / (1.0f + __expf(-omr[18 + k]));
  const float py = (float)(y - 1 + k / 3) + dy;
  const float px = (float)(x - 1 + k % 3) + dx;
  const float fy = floorf(py), fx = floorf(px);
  t.y0 = (int)fy; t.x0 = (int)fx;
  t.wy1 = py - fy; t.wx1 = px - fx;
  const float wy0 = 1.0f - t.wy1, wx0 = 1.0f - t.wx1;
  const bool yv0 = t.y0 >= 0 && t.y0 < H, yv1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  const bool xv0 = t.x0 >= 0 && t.x0 < W, xv1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  t.v00 = yv0 && xv0; t.v01 = yv0 && xv1; t.v10 = yv1 && xv0; t.v11 = yv1 && xv1;
  t.w00 = wy0 * wx0; t.w01 = wy0 * t.wx1; t.w10 = t.wy1 * wx0; t.w11 = t.wy1 * t.wx1;
  return t;
}

__global__ void __launch_bounds__(256) dcn_im2col_kernel(const float* __restrict__ x, const float* __restrict__ om, int ldom,
                                                         float* __restrict__ col, int B, int H, int W, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  const int64_t items = (int64_t)B * H * W * 9;
  const int C4 = C >> 2;
  for (int64_t it = warp; it < items; it += nwarps) {
    const int k = (int)(it % 9);
    const int64_t p = it / 9;
    const int xx = (int)(p % W), yy = (int)((p / W) % H);
    const int b = (int)(p / ((int64_t)H * W));
    const Tap t = make_tap(om + p * ldom, k, yy, xx, H, W);
    const float* xb = x + (int64_t)b * H * W * C;
    float* dst = col + p * 9 * C + (int64_t)k * C;
    for (int c = lane; c < C4; c += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      auto add = [&](bool v, int y, int xq, float w) {
        if (v) {
          const float4 s = *reinterpret_cast<const float4*>(xb + ((int64_t)y * W + xq) * C + c * 4);
          acc.x = fmaf(w, s.x, acc.x); acc.y = fmaf(w, s.y, acc.y); acc.z = fmaf(w, s.z, acc.z); acc.w = fmaf(w, s.w, acc.w);
        }
      };
      add(t.v00, t.y0, t.x0, t.w00); add(t.v01, t.y0, t.x0 + 1, t.w01);
      add(t.v10, t.y0 + 1, t.x0, t.w10); add(t.v11, t.y0 + 1, t.x0 + 1, t.w11);
      acc.x *= t.m; acc.y *= t.m; acc.z *= t.m; acc.w *= t.m;
      *reinterpret_cast<float4*>(dst + c * 4) = acc;
    }
  }
}

// adjoint: dx (atomic scatter), d(offset y/x), d(mask logit)
__global__ void __launch_bounds__(256) dcn_col2im_kernel(const float* __restrict__ x, const float* __restrict__ om, int ldom,
                                                         const float* __restrict__ dcol, float* __restrict__ dx,
                                                         float* __restrict__ dom, int B, int H, int W, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  const int64_t items = (int64_t)B * H * W * 9;
  for (int64_t it = warp; it < items; it += nwarps) {
    const int k = (int)(it % 9);
    const int64_t p = it / 9;
    const int xx = (int)(p % W), yy = (int)((p / W) % H);
    const int b = (int)(p / ((int64_t)H * W));
    const Tap t = make_tap(om + p * ldom, k, yy, xx, H, W);
    const float* xb = x + (int64_t)b * H * W * C;
    float* dxb = dx + (int64_t)b * H * W * C;
    const float* g = dcol + p * 9 * C + (int64_t)k * C;
    float gy = 0.f, gx = 0.f, gm = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float gc = g[c];
      const float s00 = t.v00 ? xb[((int64_t)t.y0 * W + t.x0) * C + c] : 0.f;
      const float s01 = t.v01 ? xb[((int64_t)t.y0 * W + t.x0 + 1) * C + c] : 0.f;
      const float s10 = t.v10 ? xb[((int64_t)(t.y0 + 1) * W + t.x0) * C + c] : 0.f;
      const float s11 = t.v11 ? xb[((int64_t)(t.y0 + 1) * W + t.x0 + 1) * C + c] : 0.f;
      const float val = t.w00 * s00 + t.w01 * s01 + t.w10 * s10 + t.w11 * s11;
      gm += gc * val;
      const float gv = gc * t.m;
      // d val / d py = (1-wx1)*(s10 - s00) + wx1*(s11 - s01) ; d val / d px similarly
      gy += gv * ((1.0f - t.wx1) * (s10 - s00) + t.wx1 * (s11 - s01));
      gx += gv * ((1.0f - t.wy1) * (s01 - s00) + t.wy1 * (s11 - s10));
      if (t.v00) atomicAdd(&dxb[((int64_t)t.y0 * W + t.x0) * C + c], gv * t.w00);
      if (t.v01) atomicAdd(&dxb[((int64_t)t.y0 * W + t.x0 + 1) * C + c], gv * t.w01);
      if (t.v10) atomicAdd(&dxb[((int64_t)(t.y0 + 1) * W + t.x0) * C + c], gv * t.w10);
      if (t.v11) atomicAdd(&dxb[((int64_t)(t.y0 + 1) * W + t.x0 + 1) * C + c], gv * t.w11);
    }
    gy = warp_sum(gy); gx = warp_sum(gx); gm = warp_sum(gm);
    if (lane == 0) {
      float* d = dom + p * ldom;
      d[2 * k] = gy;
      d[2 * k + 1] = gx;
      d[18 + k] = gm * t.m * (1.0f - t.m);
    }
  }
}

}  // namespace

extern "C" {

int fa_dcn_im2col(const float* x, const float* om, int ldom, float* col, int B, int H, int W, int C, fa_stream_t stream) {
  FA_REQUIRE(x && om && col && ldom >= 27, "fa_dcn_im2col: null pointer or ldom < 27");
  FA_REQUIRE(C % 4 == 0, "fa_dcn_im2col: C=%d must be a multiple of 4", C);
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_DCN, st);
  const int64_t items = (int64_t)B * H * W * 9;
  if (items == 0) return FA_OK;
  int64_t blocks = (items + 7) / 8;
  if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
  dcn_im2col_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, om, ldom, col, B, H, W, C);
  FA_LAUNCH_CHECK("fa_dcn_im2col");
  return FA_OK;
}

int fa_dcn_col2im(const float* x, const float* om, int ldom, const float* dcol, float* dx, float* dom, int B, int H, int W,
                  int C, fa_stream_t stream) {
  FA_REQUIRE(x && om && dcol && dx && dom && ldom >= 27, "fa_dcn_col2im: null pointer or ldom < 27");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_DCN, st);
  const int64_t items = (int64_t)B * H * W * 9;
  if (items == 0) return FA_OK;
  int64_t blocks = (items + 7) / 8;
  if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
  dcn_col2im_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, om, ldom, dcol, dx, dom, B, H, W, C);
  FA_LAUNCH_CHECK("fa_dcn_col2im");
  return FA_OK;
}

}  // extern "C"
