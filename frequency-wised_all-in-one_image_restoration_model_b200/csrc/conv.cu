// Token-layout (NHWC) convolution pieces: LeFF depthwise 3x3 with fused GELU, patch gather/scatter for
// dense convs run as GEMMs, ConvTranspose 2x2 pixel shuffle, strided copies and NCHW<->tokens transposes.
// All HBM-bound: channel-contiguous 128-bit accesses, one thread per (pixel, 4 channels).
#include "freqair_internal.h"

namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}

// ------------------------------------------------------------------ depthwise 3x3 (LeFF)
// Sliding-window strips: a thread owns one channel quad (its 9 weight float4s live in registers for the whole
// kernel) and walks a strip of DW_R rows x `seg` columns left to right, keeping a 3-column x (DW_R+2)-row window of
// float4s in registers, so each step loads DW_R+2 values for DW_R outputs (1.6 loads per output instead of 9, and no
// per-tap weight traffic).  Lanes of a warp are consecutive channel quads: every access is a contiguous 512-byte row.
// weights in the reference layout [C][1][3][3].
constexpr int DW_R = 4;
struct StripGeom { int B, H, W, C, nys, nxs, seg; };

__device__ __forceinline__ void strip_decode(int strip, const StripGeom& g, int& b, int& y0, int& x0) {
  const int xs = strip % g.nxs; strip /= g.nxs;
  const int ys = strip % g.nys;
  b = strip / g.nys;
  y0 = ys * DW_R;
  x0 = xs * g.seg;
}
__device__ __forceinline__ void strip_loadcol(const float* __restrict__ base, int x, int y0, const StripGeom& g,
                                              float4 (&col)[DW_R + 2]) {
#pragma unroll
  for (int rr = 0; rr < DW_R + 2; ++rr) {
    const int yy = y0 - 1 + rr;
    col[rr] = (x >= 0 && x < g.W && yy >= 0 && yy < g.H) ? ld4(base + ((int64_t)yy * g.W + x) * g.C)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool ok) {
  const int sz = ok ? 16 : 0;                                      // src-size 0: zero fill, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 lds4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}

// u2 = dwconv(h1) + b ; h2 = gelu(u2).  The input column x+1 of every step comes through a thread-private cp.async
// ring in shared memory (see the backward kernel below for the reasoning): DWF_RING steps of DW_R+2 float4 per
// thread stay in flight, DWF_WARPS warps per block, one persistent block per SM.
constexpr int DWF_WARPS = 12;
constexpr int DWF_RING = 4;
constexpr int DWF_SLOT = DW_R + 2;
constexpr int DWF_THREADS = DWF_WARPS * 32;
constexpr int DWF_SMEM = DWF_RING * DWF_SLOT * DWF_THREADS * 16;

__global__ void __launch_bounds__(DWF_THREADS, 1) dwconv_strip_kernel(const float* __restrict__ in,
                                                                      const float* __restrict__ w,
                                                                      const float* __restrict__ bias,
                                                                      float* __restrict__ outA, float* __restrict__ outB,
                                                                      StripGeom g, int nstrips, int a_mode) {
  extern __shared__ __align__(16) float dwf_smem[];
  const int c = (blockIdx.y * 32 + threadIdx.x) * 4;
  if (c >= g.C) return;
  constexpr uint32_t PL = DWF_THREADS * 16;                         // bytes between consecutive values of a thread
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(dwf_smem) + (threadIdx.y * 32 + threadIdx.x) * 16;
  float4 wv[9];
#pragma unroll
  for (int t = 0; t < 9; ++t)
    wv[t] = make_float4(w[(c + 0) * 9 + t], w[(c + 1) * 9 + t], w[(c + 2) * 9 + t], w[(c + 3) * 9 + t]);
  const float4 bv = bias ? ld4(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int strip = blockIdx.x * blockDim.y + threadIdx.y; strip < nstrips; strip += gridDim.x * blockDim.y) {
    int b, y0, x0;
    strip_decode(strip, g, b, y0, x0);
    const int64_t img = (int64_t)b * g.H * g.W * g.C + c;
    const float* base = in + img;
    const int x1 = min(g.W, x0 + g.seg);
    auto issue = [&](int x, int s) {
      if (x < x1) {
        const uint32_t a = ring + (uint32_t)(s * DWF_SLOT) * PL;
#pragma unroll
        for (int rr = 0; rr < DW_R + 2; ++rr) {
          const int yy = y0 - 1 + rr;
          const bool ok = x + 1 < g.W && yy >= 0 && yy < g.H;
          cp_async16(a + rr * PL, ok ? base + ((int64_t)yy * g.W + x + 1) * g.C : base, ok);
        }
      }
      cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < DWF_RING; ++s) issue(x0 + s, s);
    float4 win[3][DW_R + 2];
    strip_loadcol(base, x0 - 1, y0, g, win[1]);
    strip_loadcol(base, x0, y0, g, win[2]);
    int s = 0;
    for (int x = x0; x < x1; ++x) {
      cp_async_wait<DWF_RING - 1>();
      const uint32_t a = ring + (uint32_t)(s * DWF_SLOT) * PL;
#pragma unroll
      for (int rr = 0; rr < DW_R + 2; ++rr) { win[0][rr] = win[1][rr]; win[1][rr] = win[2][rr]; win[2][rr] = lds4(a + rr * PL); }
      issue(x + DWF_RING, s);
      s = (s + 1 == DWF_RING) ? 0 : s + 1;
#pragma unroll
      for (int r = 0; r < DW_R; ++r) {
        if (y0 + r >= g.H) break;
        float4 acc = bv;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) acc = fma4(win[kx][r + ky], wv[ky * 3 + kx], acc);
        const int64_t o = img + ((int64_t)(y0 + r) * g.W + x) * g.C;
        if (a_mode == 0) {
          if (outA) st4(outA + o, acc);
          if (outB) st4(outB + o, make_float4(gelu_f(acc.x), gelu_f(acc.y), gelu_f(acc.z), gelu_f(acc.w)));
        } else {                                   // outA = gelu'(u2): what the backward multiplies by, from the same cdf / exp
          float4 dgv, gv;
          gv.x = gelu_pair_f(acc.x, dgv.x); gv.y = gelu_pair_f(acc.y, dgv.y);
          gv.z = gelu_pair_f(acc.z, dgv.z); gv.w = gelu_pair_f(acc.w, dgv.w);
          st4(outA + o, dgv);
          if (outB) st4(outB + o, gv);
        }
      }
    }
    cp_async_wait<0>();
  }
}

// Backward of the LeFF depthwise conv in ONE sweep.  With q = p + delta(tap) the weight gradient
//   dw[c][tap] = sum_p du2[p,c] * h1[p + delta(tap), c] = sum_q du2[q - delta(tap), c] * h1[q, c]
// uses exactly the 3x3 window of du2 around q that the data gradient du1[q] = gelu'(u1[q]) * sum_tap w[tap]*du2[q-delta]
// needs, so both come from one window over du2 plus the centre values of h1 and u1: 4.5 passes over the hidden tensor
// (du2 x1.5, h1, u1, du1) instead of 6.2 for two kernels.  36 + 4 accumulators per thread persist over the strips a
// thread visits; the 8 strip lanes of a block are reduced in shared memory, one atomic per (channel, tap) per block.
//
// Memory-level parallelism comes from a THREAD-PRIVATE cp.async ring in shared memory (the first version kept every
// load in registers: 128 registers, spills, 16 warps/SM and 7.5 warps stalled on long-scoreboard per issue at 1.4
// TB/s).  Step x needs 14 float4 per thread - the du2 column x+1 (DW_R+2 rows), u1 and h1 at column x (DW_R rows
// each); a thread copies exactly the values it will read itself, so cp.async.wait_group is the only synchronisation.
// With DWB_RING steps per thread and 256 threads, ~115 KB per SM are always in flight (HBM latency x 6.5 TB/s / 148
// SMs ~ 45 KB), one persistent block per SM.
constexpr int DWB_RING = 3;
constexpr int DWB_SLOT = (DW_R + 2) + 2 * DW_R;                  // float4 per thread per step
constexpr int DWB_SMEM = DWB_RING * DWB_SLOT * 256 * 16;

__global__ void __launch_bounds__(256, 1) dwconv_bwd_strip_kernel(const float* __restrict__ du2,
                                                               const float* __restrict__ h1,
                                                               const float* __restrict__ u1,
                                                               const float* __restrict__ w, float* __restrict__ du1,
                                                               float* __restrict__ dw, float* __restrict__ db,
                                                               StripGeom g, int nstrips) {
  extern __shared__ __align__(16) float dwb_smem[];
  float (*red)[32][41] = reinterpret_cast<float (*)[32][41]>(dwb_smem);          // aliases the ring after the sweep
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(dwb_smem) + tid * 16;
  const int c = (blockIdx.y * 32 + threadIdx.x) * 4;
  const bool cok = c < g.C;
  float4 wv[9], acc[9];
  float4 accb = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    // window slot t = (ky,kx) holds du2[q + (ky-1, kx-1)] = du2[q - delta(8-t)]: it pairs with tap 8-t
    wv[t] = cok ? make_float4(w[(c + 0) * 9 + 8 - t], w[(c + 1) * 9 + 8 - t], w[(c + 2) * 9 + 8 - t], w[(c + 3) * 9 + 8 - t])
                : make_float4(0.f, 0.f, 0.f, 0.f);
    acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (cok) {
    for (int strip = blockIdx.x * blockDim.y + threadIdx.y; strip < nstrips; strip += gridDim.x * blockDim.y) {
      int b, y0, x0;
      strip_decode(strip, g, b, y0, x0);
      const int64_t img = (int64_t)b * g.H * g.W * g.C + c;
      const float* base = du2 + img;
      const int x1 = min(g.W, x0 + g.seg);
      // copies of step x into ring slot s
      auto issue = [&](int x, int s) {
        const uint32_t a = ring + (uint32_t)(s * DWB_SLOT) * 4096u;
        if (x < x1) {                                              // past the strip: an empty group keeps the count
#pragma unroll
          for (int rr = 0; rr < DW_R + 2; ++rr) {
            const int yy = y0 - 1 + rr;
            const bool ok = x + 1 < g.W && yy >= 0 && yy < g.H;
            cp_async16(a + rr * 4096u, ok ? base + ((int64_t)yy * g.W + x + 1) * g.C : base, ok);
          }
#pragma unroll
          for (int r = 0; r < DW_R; ++r) {
            const bool ok = y0 + r < g.H;
            const int64_t o = ok ? img + ((int64_t)(y0 + r) * g.W + x) * g.C : 0;
            if (u1) cp_async16(a + (DW_R + 2 + r) * 4096u, u1 + o, ok);
            if (dw && h1) cp_async16(a + (2 * DW_R + 2 + r) * 4096u, h1 + o, ok);
          }
        }
        cp_async_commit();
      };
#pragma unroll
      for (int s = 0; s < DWB_RING; ++s) issue(x0 + s, s);
      float4 win[3][DW_R + 2];
      strip_loadcol(base, x0 - 1, y0, g, win[1]);
      strip_loadcol(base, x0, y0, g, win[2]);
      int s = 0;
      for (int x = x0; x < x1; ++x) {
        cp_async_wait<DWB_RING - 1>();
        const uint32_t a = ring + (uint32_t)(s * DWB_SLOT) * 4096u;
#pragma unroll
        for (int rr = 0; rr < DW_R + 2; ++rr) { win[0][rr] = win[1][rr]; win[1][rr] = win[2][rr]; win[2][rr] = lds4(a + rr * 4096u); }
        float4 uq[DW_R], hq[DW_R];
#pragma unroll
        for (int r = 0; r < DW_R; ++r) {
          if (u1) uq[r] = lds4(a + (DW_R + 2 + r) * 4096u);
          if (dw && h1) hq[r] = lds4(a + (2 * DW_R + 2 + r) * 4096u);
        }
        issue(x + DWB_RING, s);                                   // the slot's values are in registers now
        s = (s + 1 == DWB_RING) ? 0 : s + 1;
#pragma unroll
        for (int r = 0; r < DW_R; ++r) {
          if (y0 + r >= g.H) break;
          const int64_t o = img + ((int64_t)(y0 + r) * g.W + x) * g.C;
          float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) d = fma4(win[kx][r + ky], wv[ky * 3 + kx], d);
          if (u1) {
            const float4 u = uq[r];
            float4 dg, gv;                                       // gelu'(u1) and h1 = gelu(u1) from one cdf / exp
            gv.x = gelu_pair_f(u.x, dg.x); gv.y = gelu_pair_f(u.y, dg.y);
            gv.z = gelu_pair_f(u.z, dg.z); gv.w = gelu_pair_f(u.w, dg.w);
            d = make_float4(d.x * dg.x, d.y * dg.y, d.z * dg.z, d.w * dg.w);
            if (!h1) hq[r] = gv;                                 // h1 not passed: recomputed, one stream less to read
          }
          st4(du1 + o, d);
          if (dw) {
            const float4 gc = win[1][r + 1];                       // du2[q]
            accb.x += gc.x; accb.y += gc.y; accb.z += gc.z; accb.w += gc.w;
#pragma unroll
            for (int t = 0; t < 9; ++t) acc[t] = fma4(win[t % 3][r + t / 3], hq[r], acc[t]);
          }
        }
      }
      cp_async_wait<0>();                                          // only zero-size tail copies are left
    }
  }
  if (!dw) return;                                                  // uniform over the block
  __syncthreads();                                                  // every thread is done with its ring slots
  const int pl = threadIdx.y, cq = threadIdx.x;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    // acc[t] belongs to tap 8-t
    red[pl][cq][(8 - t) * 4 + 0] = acc[t].x; red[pl][cq][(8 - t) * 4 + 1] = acc[t].y;
    red[pl][cq][(8 - t) * 4 + 2] = acc[t].z; red[pl][cq][(8 - t) * 4 + 3] = acc[t].w;
  }
  red[pl][cq][36] = accb.x; red[pl][cq][37] = accb.y; red[pl][cq][38] = accb.z; red[pl][cq][39] = accb.w;
  __syncthreads();
  for (int i = tid; i < 32 * 40; i += 256) {
    const int q = i / 40, v = i % 40;
    const int cc = (blockIdx.y * 32 + q) * 4;
    if (cc >= g.C) continue;
    float sum = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) sum += red[l][q][v];
    if (v < 36) atomicAdd(&dw[(cc + (v & 3)) * 9 + (v >> 2)], sum);
    else if (db) atomicAdd(&db[cc + (v - 36)], sum);
  }
}

// ------------------------------------------------------------------ 3x3 conv to a few output channels (OutputProj)
// tokens [B,H*W,C] -> NCHW image [B,CO,H*W] (+ bias + residual image), CO <= 4, C <= 128 (decoder_Uformer.py:476-499,
// :1171).  A 1008-wide patch matrix for 3 output channels is 1 GB of pure traffic per direction at B=16; here a warp
// owns a pixel, lane = channel quad with its 9 x CO weight float4s in registers, and the three dot products are warp
// reductions: the feature map is read once (through L1 for the 9 taps) and nothing is materialised.
// wk layout: [CO][(ky,kx,ci)] - the GEMM weight matrix of convs.conv_weight_matrix.
template <int CO>
__global__ void __launch_bounds__(256) conv3x3_out_fwd_kernel(const float* __restrict__ t, const float* __restrict__ wk,
                                                              const float* __restrict__ bias,
                                                              const float* __restrict__ ximg, float* __restrict__ out,
                                                              int B, int H, int W, int C) {
  const int lane = threadIdx.x & 31;
  const int c = lane * 4;
  const bool cok = c < C;
  float4 w[CO][9];
#pragma unroll
  for (int co = 0; co < CO; ++co)
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) w[co][tp] = cok ? ld4(wk + ((int64_t)co * 9 + tp) * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int HW = H * W;
  const int64_t total = (int64_t)B * HW;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t p = warp; p < total; p += nwarps) {
    const int b = (int)(p / HW), pix = (int)(p % HW), y = pix / W, x = pix % W;
    float acc[CO];
#pragma unroll
    for (int co = 0; co < CO; ++co) acc[co] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1;
        if (xx < 0 || xx >= W || !cok) continue;
        const float4 v = ld4(t + (((int64_t)b * H + yy) * W + xx) * C + c);
#pragma unroll
        for (int co = 0; co < CO; ++co) {
          const float4 ww = w[co][ky * 3 + kx];
          acc[co] = fmaf(v.x, ww.x, fmaf(v.y, ww.y, fmaf(v.z, ww.z, fmaf(v.w, ww.w, acc[co]))));
        }
      }
    }
#pragma unroll
    for (int co = 0; co < CO; ++co) {
      const float sum = warp_sum(acc[co]);
      if (lane == co) {
        const int64_t o = ((int64_t)b * CO + co) * HW + pix;
        out[o] = sum + (bias ? bias[co] : 0.f) + (ximg ? ximg[o] : 0.f);
      }
    }
  }
}

// dt[p][ci] = sum_tap sum_co dout[p - delta(tap)][co] * w[co][tap][ci]
template <int CO>
__global__ void __launch_bounds__(256) conv3x3_out_bwd_data_kernel(const float* __restrict__ dout,
                                                                   const float* __restrict__ wk, float* __restrict__ dt,
                                                                   int B, int H, int W, int C) {
  const int lane = threadIdx.x & 31;
  const int c = lane * 4;
  const bool cok = c < C;
  float4 w[CO][9];
#pragma unroll
  for (int co = 0; co < CO; ++co)
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) w[co][tp] = cok ? ld4(wk + ((int64_t)co * 9 + tp) * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int HW = H * W;
  const int64_t total = (int64_t)B * HW;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t p = warp; p < total; p += nwarps) {
    const int b = (int)(p / HW), pix = (int)(p % HW), y = pix / W, x = pix % W;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y - (ky - 1);
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x - (kx - 1);
        if (xx < 0 || xx >= W) continue;
#pragma unroll
        for (int co = 0; co < CO; ++co) {
          const float g = __ldg(dout + ((int64_t)b * CO + co) * HW + yy * W + xx);       // warp-uniform address
          const float4 ww = w[co][ky * 3 + kx];
          acc = fma4(make_float4(g, g, g, g), ww, acc);
        }
      }
    }
    if (cok) st4(dt + p * C + c, acc);
  }
}

// dW[co][tap][ci] += sum_p dout[p][co] * t[p + delta(tap)][ci];  db[co] += sum_p dout[p][co]
template <int CO>
__global__ void __launch_bounds__(256) conv3x3_out_bwd_weight_kernel(const float* __restrict__ t,
                                                                     const float* __restrict__ dout,
                                                                     float* __restrict__ dW, float* __restrict__ db,
                                                                     int B, int H, int W, int C) {
  __shared__ float red[32][CO * 9 * 4 + 1];
  __shared__ float redb[CO];
  const int lane = threadIdx.x & 31;
  const int c = lane * 4;
  const bool cok = c < C;
  for (int i = threadIdx.x; i < 32 * (CO * 9 * 4 + 1); i += blockDim.x) (&red[0][0])[i] = 0.f;
  if (threadIdx.x < CO) redb[threadIdx.x] = 0.f;
  __syncthreads();
  float4 acc[CO][9];
  float accb[CO];
#pragma unroll
  for (int co = 0; co < CO; ++co) {
    accb[co] = 0.f;
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) acc[co][tp] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int HW = H * W;
  const int64_t total = (int64_t)B * HW;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t p = warp; p < total; p += nwarps) {
    const int b = (int)(p / HW), pix = (int)(p % HW), y = pix / W, x = pix % W;
    float g[CO];
#pragma unroll
    for (int co = 0; co < CO; ++co) { g[co] = __ldg(dout + ((int64_t)b * CO + co) * HW + pix); accb[co] += g[co]; }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1;
        if (xx < 0 || xx >= W || !cok) continue;
        const float4 v = ld4(t + (((int64_t)b * H + yy) * W + xx) * C + c);
#pragma unroll
        for (int co = 0; co < CO; ++co) acc[co][ky * 3 + kx] = fma4(make_float4(g[co], g[co], g[co], g[co]), v, acc[co][ky * 3 + kx]);
      }
    }
  }
#pragma unroll
  for (int co = 0; co < CO; ++co)
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      float* r = &red[lane][(co * 9 + tp) * 4];
      atomicAdd(r + 0, acc[co][tp].x); atomicAdd(r + 1, acc[co][tp].y);
      atomicAdd(r + 2, acc[co][tp].z); atomicAdd(r + 3, acc[co][tp].w);
    }
  if (lane == 0) {
#pragma unroll
    for (int co = 0; co < CO; ++co) atomicAdd(&redb[co], accb[co]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * CO * 9 * 4; i += blockDim.x) {
    const int l = i / (CO * 9 * 4), v = i % (CO * 9 * 4);
    const int cc = l * 4 + (v & 3), ct = v >> 2;              // ct = co*9 + tap
    if (cc < C) atomicAdd(&dW[(int64_t)ct * C + cc], red[l][v]);
  }
  if (db && threadIdx.x < CO) atomicAdd(&db[threadIdx.x], redb[threadIdx.x]);
}

// ------------------------------------------------------------------ im2col / col2im
template <bool NCHW>
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ x, float* __restrict__ col, int B, int H,
                                                     int W, int C, int kh, int kw, int stride, int pad, int Ho, int Wo) {
  const int K = kh * kw * C;
  const int64_t total = (int64_t)B * Ho * Wo * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    int64_t r = i / K;
    const int ci = k % C;
    const int kk = k / C;
    const int kx = kk % kw, ky = kk / kw;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = NCHW ? x[(((int64_t)b * C + ci) * H + iy) * W + ix] : x[(((int64_t)b * H + iy) * W + ix) * C + ci];
    col[i] = v;
  }
}

__global__ void __launch_bounds__(256) im2col_vec_kernel(const float* __restrict__ x, float* __restrict__ col, int B,
                                                         int H, int W, int C, int kh, int kw, int stride, int pad,
                                                         int Ho, int Wo) {
  const int C4 = C >> 2;
  const int K4 = kh * kw * C4;
  const int64_t total = (int64_t)B * Ho * Wo * K4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K4);
    int64_t r = i / K4;
    const int ci = (k % C4) * 4;
    const int kk = k / C4;
    const int kx = kk % kw, ky = kk / kw;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = ld4(x + (((int64_t)b * H + iy) * W + ix) * C + ci);
    st4(col + i * 4, v);
  }
}

__global__ void __launch_bounds__(256) col2im_kernel(const float* __restrict__ col, float* __restrict__ dx, int B, int H,
                                                     int W, int C, int kh, int kw, int stride, int pad, int Ho, int Wo) {
  const int K = kh * kw * C;
  const int64_t total = (int64_t)B * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % C);
    int64_t r = i / C;
    const int ix = (int)(r % W); r /= W;
    const int iy = (int)(r % H);
    const int b = (int)(r / H);
    float s = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
      const int ty = iy + pad - ky;
      if (ty < 0 || ty % stride) continue;
      const int oy = ty / stride;
      if (oy >= Ho) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int tx = ix + pad - kx;
        if (tx < 0 || tx % stride) continue;
        const int ox = tx / stride;
        if (ox >= Wo) continue;
        s += col[(((int64_t)b * Ho + oy) * Wo + ox) * K + (ky * kw + kx) * C + ci];
      }
    }
    dx[i] = s;
  }
}

// ------------------------------------------------------------------ ConvTranspose 2x2 s2 scatter/gather
__global__ void __launch_bounds__(256) pixshuf_kernel(const float* __restrict__ g, float* __restrict__ y, int64_t ldy,
                                                      int B, int H, int W, int Co, bool fwd) {
  const int C4 = Co >> 2;
  const int64_t total = (int64_t)B * H * W * 4 * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    int64_t r = i / C4;
    const int kk = (int)(r % 4); r /= 4;
    const int x = (int)(r % W); r /= W;
    const int yy = (int)(r % H);
    const int b = (int)(r / H);
    const int ky = kk >> 1, kx = kk & 1;
    const int64_t gi = ((((int64_t)b * H + yy) * W + x) * 4 + kk) * Co + c;
    const int64_t yi = (((int64_t)b * 2 * H + 2 * yy + ky) * 2 * W + 2 * x + kx) * ldy + c;
    if (fwd) st4(y + yi, ld4(g + gi)); else st4(const_cast<float*>(g) + gi, ld4(y + yi));
  }
}

// ------------------------------------------------------------------ strided copy / add
__global__ void __launch_bounds__(256) copy2d_kernel(const float* __restrict__ a, int64_t lda,
                                                     const float* __restrict__ b, int64_t ldb, float* __restrict__ d,
                                                     int64_t ldd, int64_t rows, int cols,
                                                     const float* __restrict__ rowscale, int rps) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i % cols);
    float v = a[r * lda + c];
    if (rowscale) v *= rowscale[r / rps];
    if (b) v += b[r * ldb + c];
    d[r * ldd + c] = v;
  }
}

// ------------------------------------------------------------------ [B][R][Cc] -> [B][Cc][R] transpose (+ residual)
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, const float* __restrict__ res,
                                                        float* __restrict__ out, int R, int Cc) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* ib = in + (int64_t)b * R * Cc;
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < R && c < Cc) ? ib[(int64_t)r * Cc + c] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (c < Cc && r < R) {
      const int64_t o = (int64_t)b * R * Cc + (int64_t)c * R + r;
      out[o] = tile[tx][j] + (res ? res[o] : 0.f);
    }
  }
}

inline int ew_grid(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  if (blocks > cap) blocks = cap;
  return (int)(blocks > 0 ? blocks : 1);
}

}  // namespace

template <int CO>
static int conv_out_launch(int what, const float* t, const float* wk, const float* bias, const float* ximg, float* out,
                           const float* dout, float* dt, float* dW, float* db, int B, int H, int W, int C,
                           cudaStream_t st) {
  const int64_t pixels = (int64_t)B * H * W;
  int64_t blocks = (pixels + 63) / 64;                       // >= 8 pixels per warp
  const int64_t cap = (int64_t)kNumSMs * 8;
  const int grid = (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
  if (what == 0) conv3x3_out_fwd_kernel<CO><<<grid, 256, 0, st>>>(t, wk, bias, ximg, out, B, H, W, C);
  else if (what == 1) conv3x3_out_bwd_data_kernel<CO><<<grid, 256, 0, st>>>(dout, wk, dt, B, H, W, C);
  else conv3x3_out_bwd_weight_kernel<CO><<<grid < 2 * kNumSMs ? grid : 2 * kNumSMs, 256, 0, st>>>(t, dout, dW, db, B, H, W, C);
  return FA_OK;
}

extern "C" {

// One persistent block per SM for the ring kernels: pick the strip length (32, 16 or 8 columns) that minimises
// rounds x (strip length + halo / pipeline-fill cost) - every warp of every block walks `rounds` strips (wave
// quantisation was 13 % of the first version's time, and the 32^2...8^2 levels have too few long strips for 148 SMs).
static StripGeom make_strips_persistent(int B, int H, int W, int C, int warps, int& nstrips, dim3& grid) {
  const int gy = (C / 4 + 31) / 32;
  const int max_gx = kNumSMs / gy > 0 ? kNumSMs / gy : 1;
  StripGeom best{}; double best_cost = 1e30; int best_gx = 1, best_n = 0;
  for (int seg = 32; seg >= 8; seg >>= 1) {
    StripGeom g;
    g.B = B; g.H = H; g.W = W; g.C = C;
    g.seg = W < seg ? W : seg;
    g.nys = (H + DW_R - 1) / DW_R;
    g.nxs = (W + g.seg - 1) / g.seg;
    const int n = B * g.nys * g.nxs;
    const int nb = (n + warps - 1) / warps;                      // blocks' worth of strips
    const int rounds = (nb + max_gx - 1) / max_gx;
    const int gx = (nb + rounds - 1) / rounds;
    const double cost = (double)rounds * (g.seg + 2.5);
    if (cost < best_cost) { best_cost = cost; best = g; best_gx = gx; best_n = n; }
    if (W <= seg) break;
  }
  nstrips = best_n;
  grid = dim3((unsigned)best_gx, (unsigned)gy);
  return best;
}

int fa_dwconv3x3_fwd(const float* h1, const float* w, const float* b, float* u2, float* h2, int u2_mode, int B, int H,
                     int W, int C, fa_stream_t stream) {
  FA_REQUIRE(h1 && w && (u2 || h2), "fa_dwconv3x3_fwd: null pointer");
  FA_REQUIRE(u2_mode == 0 || (u2_mode == 1 && u2), "fa_dwconv3x3_fwd: u2_mode must be 0, or 1 with u2 set");
  FA_REQUIRE(C % 4 == 0, "fa_dwconv3x3_fwd: C=%d must be a multiple of 4", C);
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_DWCONV, st);
  if ((int64_t)B * H * W * C == 0) return FA_OK;
  int nstrips; dim3 grid;
  const StripGeom g = make_strips_persistent(B, H, W, C, DWF_WARPS, nstrips, grid);
  FA_SMEM_ATTR_ONCE(DWF_SMEM, dwconv_strip_kernel);
  dwconv_strip_kernel<<<grid, dim3(32, DWF_WARPS), DWF_SMEM, st>>>(h1, w, b, u2, h2, g, nstrips, u2_mode);
  FA_LAUNCH_CHECK("fa_dwconv3x3_fwd");
  return FA_OK;
}

int fa_dwconv3x3_bwd(const float* du2, const float* h1, const float* u1, const float* w, float* du1, float* dw,
                     float* db, int B, int H, int W, int C, fa_stream_t stream) {
  FA_REQUIRE(du2 && w && du1, "fa_dwconv3x3_bwd: null pointer");
  FA_REQUIRE(C % 4 == 0, "fa_dwconv3x3_bwd: C=%d must be a multiple of 4", C);
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_DWCONV, st);
  if ((int64_t)B * H * W * C == 0) return FA_OK;
  int nstrips; dim3 grid;
  const StripGeom g = make_strips_persistent(B, H, W, C, 8, nstrips, grid);
  FA_REQUIRE(!dw || h1 || u1, "fa_dwconv3x3_bwd: the weight gradient needs h1, or u1 to recompute h1 = gelu(u1)");
  FA_SMEM_ATTR_ONCE(DWB_SMEM, dwconv_bwd_strip_kernel);
  dwconv_bwd_strip_kernel<<<grid, dim3(32, 8), DWB_SMEM, st>>>(du2, h1, u1, w, du1, dw, db, g, nstrips);
  FA_LAUNCH_CHECK("fa_dwconv3x3_bwd");
  return FA_OK;
}

int fa_conv3x3_out_fwd(const float* t, const float* wk, const float* bias, const float* ximg, float* out, int B, int H,
                       int W, int C, int Co, fa_stream_t stream) {
  FA_REQUIRE(t && wk && out, "fa_conv3x3_out_fwd: null pointer");
  FA_REQUIRE(C % 4 == 0 && C <= 128 && Co >= 1 && Co <= 4, "fa_conv3x3_out_fwd: C=%d (mult of 4, <=128), Co=%d (1..4)", C, Co);
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_IM2COL, st);
  if ((int64_t)B * H * W == 0) return FA_OK;
  switch (Co) {
    case 1: conv_out_launch<1>(0, t, wk, bias, ximg, out, nullptr, nullptr, nullptr, nullptr, B, H, W, C, st); break;
    case 2: conv_out_launch<2>(0, t, wk, bias, ximg, out, nullptr, nullptr, nullptr, nullptr, B, H, W, C, st); break;
    case 3: conv_out_launch<3>(0, t, wk, bias, ximg, out, nullptr, nullptr, nullptr, nullptr, B, H, W, C, st); break;
    default: conv_out_launch<4>(0, t, wk, bias, ximg, out, nullptr, nullptr, nullptr, nullptr, B, H, W, C, st); break;
  }
  FA_LAUNCH_CHECK("fa_conv3x3_out_fwd");
  return FA_OK;
}

int fa_conv3x3_out_bwd(const float* t, const float* wk, const float* dout, float* dt, float* dW, float* db, int B, int H,
                       int W, int C, int Co, fa_stream_t stream) {
  FA_REQUIRE(t && wk && dout, "fa_conv3x3_out_bwd: null pointer");
  FA_REQUIRE(C % 4 == 0 && C <= 128 && Co >= 1 && Co <= 4, "fa_conv3x3_out_bwd: C=%d (mult of 4, <=128), Co=%d (1..4)", C, Co);
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_IM2COL, st);
  if ((int64_t)B * H * W == 0) return FA_OK;
  for (int what = 1; what <= 2; ++what) {
    if (what == 1 && !dt) continue;
    if (what == 2 && !dW) continue;
    if (what == 2) fa_count_launch(FA_K_IM2COL);
    switch (Co) {
      case 1: conv_out_launch<1>(what, t, wk, nullptr, nullptr, nullptr, dout, dt, dW, db, B, H, W, C, st); break;
      case 2: conv_out_launch<2>(what, t, wk, nullptr, nullptr, nullptr, dout, dt, dW, db, B, H, W, C, st); break;
      case 3: conv_out_launch<3>(what, t, wk, nullptr, nullptr, nullptr, dout, dt, dW, db, B, H, W, C, st); break;
      default: conv_out_launch<4>(what, t, wk, nullptr, nullptr, nullptr, dout, dt, dW, db, B, H, W, C, st); break;
    }
    FA_LAUNCH_CHECK("fa_conv3x3_out_bwd");
  }
  return FA_OK;
}

int fa_im2col(const float* x, float* col, int B, int H, int W, int C, int kh, int kw, int stride, int pad, int nchw_in,
              fa_stream_t stream) {
  FA_REQUIRE(x && col && stride >= 1 && kh >= 1 && kw >= 1, "fa_im2col: bad argument");
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  FA_REQUIRE(Ho > 0 && Wo > 0, "fa_im2col: empty output");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_IM2COL, st);
  const int64_t total = (int64_t)B * Ho * Wo * kh * kw * C;
  if (total == 0) return FA_OK;
  const bool al = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(col) % 16 == 0);
  if (nchw_in) im2col_kernel<true><<<ew_grid(total), 256, 0, st>>>(x, col, B, H, W, C, kh, kw, stride, pad, Ho, Wo);
  else if (C % 4 == 0 && al) im2col_vec_kernel<<<ew_grid(total / 4), 256, 0, st>>>(x, col, B, H, W, C, kh, kw, stride, pad, Ho, Wo);
  else im2col_kernel<false><<<ew_grid(total), 256, 0, st>>>(x, col, B, H, W, C, kh, kw, stride, pad, Ho, Wo);
  FA_LAUNCH_CHECK("fa_im2col");
  return FA_OK;
}

int fa_col2im(const float* col, float* dx, int B, int H, int W, int C, int kh, int kw, int stride, int pad,
              fa_stream_t stream) {
  FA_REQUIRE(col && dx && stride >= 1, "fa_col2im: bad argument");
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_IM2COL, st);
  const int64_t total = (int64_t)B * H * W * C;
  if (total == 0) return FA_OK;
  col2im_kernel<<<ew_grid(total), 256, 0, st>>>(col, dx, B, H, W, C, kh, kw, stride, pad, Ho, Wo);
  FA_LAUNCH_CHECK("fa_col2im");
  return FA_OK;
}

int fa_pixel_shuffle2_fwd(const float* g, float* y, int64_t ldy, int B, int H, int W, int Co, fa_stream_t stream) {
  FA_REQUIRE(g && y && Co % 4 == 0 && ldy % 4 == 0, "fa_pixel_shuffle2_fwd: bad argument (Co, ldy multiples of 4)");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  const int64_t total = (int64_t)B * H * W * Co;
  if (total == 0) return FA_OK;
  pixshuf_kernel<<<ew_grid(total), 256, 0, st>>>(g, y, ldy, B, H, W, Co, true);
  FA_LAUNCH_CHECK("fa_pixel_shuffle2_fwd");
  return FA_OK;
}

int fa_pixel_shuffle2_bwd(const float* dy, int64_t ldy, float* dg, int B, int H, int W, int Co, fa_stream_t stream) {
  FA_REQUIRE(dy && dg && Co % 4 == 0 && ldy % 4 == 0, "fa_pixel_shuffle2_bwd: bad argument (Co, ldy multiples of 4)");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  const int64_t total = (int64_t)B * H * W * Co;
  if (total == 0) return FA_OK;
  pixshuf_kernel<<<ew_grid(total), 256, 0, st>>>(dg, const_cast<float*>(dy), ldy, B, H, W, Co, false);
  FA_LAUNCH_CHECK("fa_pixel_shuffle2_bwd");
  return FA_OK;
}

int fa_copy2d(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int cols, fa_stream_t stream) {
  FA_REQUIRE(src && dst, "fa_copy2d: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  if (rows * cols == 0) return FA_OK;
  copy2d_kernel<<<ew_grid(rows * cols), 256, 0, st>>>(src, lds, nullptr, 0, dst, ldd, rows, cols, nullptr, 1);
  FA_LAUNCH_CHECK("fa_copy2d");
  return FA_OK;
}

int fa_add2d(const float* a, int64_t lda, const float* b, int64_t ldb, float* dst, int64_t ldd, int64_t rows, int cols,
             fa_stream_t stream) {
  FA_REQUIRE(a && b && dst, "fa_add2d: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  if (rows * cols == 0) return FA_OK;
  copy2d_kernel<<<ew_grid(rows * cols), 256, 0, st>>>(a, lda, b, ldb, dst, ldd, rows, cols, nullptr, 1);
  FA_LAUNCH_CHECK("fa_add2d");
  return FA_OK;
}

int fa_scale_rows(const float* src, const float* rowscale, int rows_per_scale, float* dst, int64_t rows, int cols,
                  fa_stream_t stream) {
  FA_REQUIRE(src && dst && rowscale && rows_per_scale > 0, "fa_scale_rows: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  if (rows * cols == 0) return FA_OK;
  copy2d_kernel<<<ew_grid(rows * cols), 256, 0, st>>>(src, cols, nullptr, 0, dst, cols, rows, cols, rowscale, rows_per_scale);
  FA_LAUNCH_CHECK("fa_scale_rows");
  return FA_OK;
}

int fa_tokens_to_nchw(const float* t, const float* res, float* y, int B, int HW, int C, fa_stream_t stream) {
  FA_REQUIRE(t && y, "fa_tokens_to_nchw: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  if ((int64_t)B * HW * C == 0) return FA_OK;
  transpose_kernel<<<dim3((C + 31) / 32, (HW + 31) / 32, B), 256, 0, st>>>(t, res, y, HW, C);
  FA_LAUNCH_CHECK("fa_tokens_to_nchw");
  return FA_OK;
}

int fa_nchw_to_tokens(const float* x, float* t, int B, int HW, int C, fa_stream_t stream) {
  FA_REQUIRE(x && t, "fa_nchw_to_tokens: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  if ((int64_t)B * HW * C == 0) return FA_OK;
  transpose_kernel<<<dim3((HW + 31) / 32, (C + 31) / 32, B), 256, 0, st>>>(x, nullptr, t, C, HW);
  FA_LAUNCH_CHECK("fa_nchw_to_tokens");
  return FA_OK;
}

}  // extern "C"
