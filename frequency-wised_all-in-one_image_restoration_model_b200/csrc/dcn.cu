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

// ---- pixel-per-warp kernels (C = 64: every DCN layer of the 64-channel DGRN) ------------------------------------------
// A warp owns one pixel at a time.  Lanes 0..8 turn the pixel's 27 offset / mask values into the 9 taps' corner offsets
// and weights ONCE and park them in shared memory; then the col row of the pixel - 9 taps x 64 channels = 2304
// contiguous bytes - is produced as 144 float4 items (tap = i / 16, channel quad = i % 16) spread over the 32 lanes:
// every global store is a fully coalesced 512-byte warp store, every corner read a 256-byte segment shared by a
// half-warp, and the 20 corner loads of a lane are independent (all in flight together).  The first version gave one
// (pixel, tap) to a warp: at C = 64 half of its lanes idled, each of its 9 warps recomputed the tap, and it ran at
// 0.75 TB/s of its 600 MB output; warps of a block take adjacent pixels so the gathers of one block hit the same rows
// of x in L1.
constexpr int PW_TAPW = 12;                    // words per tap in shared memory

__device__ __forceinline__ void tap_to_smem(const Tap& t, int W, int C, float* dstf, bool masked) {
  int* dsti = reinterpret_cast<int*>(dstf);
  dsti[0] = t.v00 ? (t.y0 * W + t.x0) * C : 0;
  dsti[1] = t.v01 ? (t.y0 * W + t.x0 + 1) * C : 0;
  dsti[2] = t.v10 ? ((t.y0 + 1) * W + t.x0) * C : 0;
  dsti[3] = t.v11 ? ((t.y0 + 1) * W + t.x0 + 1) * C : 0;
  const float m = masked ? t.m : 1.0f;         // forward: mask folded into the weights; invalid corners weigh 0
  dstf[4] = t.v00 ? t.w00 * m : 0.f; dstf[5] = t.v01 ? t.w01 * m : 0.f;
  dstf[6] = t.v10 ? t.w10 * m : 0.f; dstf[7] = t.v11 ? t.w11 * m : 0.f;
  dstf[8] = t.wy1; dstf[9] = t.wx1; dstf[10] = t.m;
  dsti[11] = (t.v00 ? 1 : 0) | (t.v01 ? 2 : 0) | (t.v10 ? 4 : 0) | (t.v11 ? 8 : 0);
}

__global__ void __launch_bounds__(256) dcn_im2col_c64_kernel(const float* __restrict__ x, const float* __restrict__ om,
                                                             int ldom, float* __restrict__ col, int B, int H, int W) {
  constexpr int C = 64;
  __shared__ __align__(16) float taps[8][9 * PW_TAPW];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* tp = taps[wib];
  const int64_t npix = (int64_t)B * H * W;
  for (int64_t p = (int64_t)blockIdx.x * 8 + wib; p < npix; p += (int64_t)gridDim.x * 8) {
    const int xx = (int)(p % W), yy = (int)((p / W) % H);
    const int64_t b = p / ((int64_t)H * W);
    __syncwarp();
    if (lane < 9) tap_to_smem(make_tap(om + p * ldom, lane, yy, xx, H, W), W, C, tp + lane * PW_TAPW, true);
    __syncwarp();
    const float* xb = x + b * H * W * C;
    float4* dst = reinterpret_cast<float4*>(col + p * 9 * C);
    float4 s[5][4];
    float4 wv[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int i = lane + 32 * j;
      if (i < 144) {
        const int k = i >> 4, c = (i & 15) * 4;
        const int4 off = *reinterpret_cast<const int4*>(tp + k * PW_TAPW);
        wv[j] = *reinterpret_cast<const float4*>(tp + k * PW_TAPW + 4);
        s[j][0] = __ldg(reinterpret_cast<const float4*>(xb + off.x + c));
        s[j][1] = __ldg(reinterpret_cast<const float4*>(xb + off.y + c));
        s[j][2] = __ldg(reinterpret_cast<const float4*>(xb + off.z + c));
        s[j][3] = __ldg(reinterpret_cast<const float4*>(xb + off.w + c));
      }
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int i = lane + 32 * j;
      if (i < 144) {
        float4 a;
        a.x = wv[j].x * s[j][0].x + wv[j].y * s[j][1].x + wv[j].z * s[j][2].x + wv[j].w * s[j][3].x;
        a.y = wv[j].x * s[j][0].y + wv[j].y * s[j][1].y + wv[j].z * s[j][2].y + wv[j].w * s[j][3].y;
        a.z = wv[j].x * s[j][0].z + wv[j].y * s[j][1].z + wv[j].z * s[j][2].z + wv[j].w * s[j][3].z;
        a.w = wv[j].x * s[j][0].w + wv[j].y * s[j][1].w + wv[j].z * s[j][2].w + wv[j].w * s[j][3].w;
        dst[i] = a;
      }
    }
  }
}

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// adjoint, same mapping: a lane takes the float4 items (tap, channel quad) of the pixel's dcol row, scatters the data
// gradient with ONE 128-bit reduction per corner (red.global.add.v4.f32: 4x fewer atomic instructions than per-channel
// atomicAdd), and the offset / mask gradients are reduced over the 16 lanes that share a tap.
__global__ void __launch_bounds__(256) dcn_col2im_c64_kernel(const float* __restrict__ x, const float* __restrict__ om,
                                                             int ldom, const float* __restrict__ dcol,
                                                             float* __restrict__ dx, float* __restrict__ dom, int B, int H,
                                                             int W) {
  constexpr int C = 64;
  __shared__ __align__(16) float taps[8][9 * PW_TAPW];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* tp = taps[wib];
  const int64_t npix = (int64_t)B * H * W;
  for (int64_t p = (int64_t)blockIdx.x * 8 + wib; p < npix; p += (int64_t)gridDim.x * 8) {
    const int xx = (int)(p % W), yy = (int)((p / W) % H);
    const int64_t b = p / ((int64_t)H * W);
    __syncwarp();
    if (lane < 9) tap_to_smem(make_tap(om + p * ldom, lane, yy, xx, H, W), W, C, tp + lane * PW_TAPW, false);
    __syncwarp();
    const float* xb = x + b * H * W * C;
    float* dxb = dx + b * H * W * C;
    const float4* g4 = reinterpret_cast<const float4*>(dcol + p * 9 * C);
    float* d = dom + p * ldom;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int i = lane + 32 * j;
      float gy = 0.f, gx = 0.f, gm = 0.f;
      int k = 0;
      if (i < 144) {
        k = i >> 4;
        const int c = (i & 15) * 4;
        const float* tk = tp + k * PW_TAPW;
        const int4 off = *reinterpret_cast<const int4*>(tk);
        const float4 w = *reinterpret_cast<const float4*>(tk + 4);       // corner weights, 0 where the corner is outside
        const float wy1 = tk[8], wx1 = tk[9], m = tk[10];
        const int vb = reinterpret_cast<const int*>(tk)[11];
        const float4 gc = g4[i];
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 s00 = (vb & 1) ? __ldg(reinterpret_cast<const float4*>(xb + off.x + c)) : z4;
        const float4 s01 = (vb & 2) ? __ldg(reinterpret_cast<const float4*>(xb + off.y + c)) : z4;
        const float4 s10 = (vb & 4) ? __ldg(reinterpret_cast<const float4*>(xb + off.z + c)) : z4;
        const float4 s11 = (vb & 8) ? __ldg(reinterpret_cast<const float4*>(xb + off.w + c)) : z4;
        const float4 gv = make_float4(gc.x * m, gc.y * m, gc.z * m, gc.w * m);
#define FA_DCN_CH(f)                                                                                        \
        {                                                                                                     \
          const float val = w.x * s00.f + w.y * s01.f + w.z * s10.f + w.w * s11.f;                            \
          gm += gc.f * val;                                                                                   \
          gy += gv.f * ((1.0f - wx1) * (s10.f - s00.f) + wx1 * (s11.f - s01.f));                              \
          gx += gv.f * ((1.0f - wy1) * (s01.f - s00.f) + wy1 * (s11.f - s10.f));                              \
        }
        FA_DCN_CH(x) FA_DCN_CH(y) FA_DCN_CH(z) FA_DCN_CH(w)
#undef FA_DCN_CH
        if (vb & 1) red_add_v4(dxb + off.x + c, make_float4(gv.x * w.x, gv.y * w.x, gv.z * w.x, gv.w * w.x));
        if (vb & 2) red_add_v4(dxb + off.y + c, make_float4(gv.x * w.y, gv.y * w.y, gv.z * w.y, gv.w * w.y));
        if (vb & 4) red_add_v4(dxb + off.z + c, make_float4(gv.x * w.z, gv.y * w.z, gv.z * w.z, gv.w * w.z));
        if (vb & 8) red_add_v4(dxb + off.w + c, make_float4(gv.x * w.w, gv.y * w.w, gv.z * w.w, gv.w * w.w));
        gm *= m * (1.0f - m);
      }
      // the 16 lanes of a half-warp share the tap: xor-shuffles over 8, 4, 2, 1 stay inside the half
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        gy += __shfl_xor_sync(0xffffffffu, gy, o);
        gx += __shfl_xor_sync(0xffffffffu, gx, o);
        gm += __shfl_xor_sync(0xffffffffu, gm, o);
      }
      if ((lane & 15) == 0 && i < 144) {
        d[2 * k] = gy;
        d[2 * k + 1] = gx;
        d[18 + k] = gm;
      }
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
  if (C == 64 && (((uintptr_t)x | (uintptr_t)col) % 16) == 0 && (int64_t)H * W * C < (1ll << 31)) {
    int64_t pb = ((int64_t)B * H * W + 7) / 8;
    if (pb > (int64_t)kNumSMs * 8) pb = (int64_t)kNumSMs * 8;
    dcn_im2col_c64_kernel<<<(unsigned)pb, 256, 0, st>>>(x, om, ldom, col, B, H, W);
  } else {
    dcn_im2col_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, om, ldom, col, B, H, W, C);
  }
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
  if (C == 64 && (((uintptr_t)x | (uintptr_t)dcol | (uintptr_t)dx) % 16) == 0 && (int64_t)H * W * C < (1ll << 31)) {
    int64_t pb = ((int64_t)B * H * W + 7) / 8;
    if (pb > (int64_t)kNumSMs * 8) pb = (int64_t)kNumSMs * 8;
    dcn_col2im_c64_kernel<<<(unsigned)pb, 256, 0, st>>>(x, om, ldom, dcol, dx, dom, B, H, W);
  } else {
    dcn_col2im_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, om, ldom, dcol, dx, dom, B, H, W, C);
  }
  FA_LAUNCH_CHECK("fa_dcn_col2im");
  return FA_OK;
}

}  // extern "C"
