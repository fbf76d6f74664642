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
// weights in the reference layout [C][1][3][3]; staged transposed as wt[tap][C] so a channel quad is one float4.
__global__ void __launch_bounds__(256) dwconv_fwd_kernel(const float* __restrict__ h1, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ u2,
                                                         float* __restrict__ h2, int B, int H, int W, int C) {
  const int C4 = C >> 2;
  const int64_t total = (int64_t)B * H * W * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    int64_t p = i / C4;
    const int x = (int)(p % W); p /= W;
    const int y = (int)(p % H);
    const int b = (int)(p / H);
    float4 acc = bias ? ld4(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1;
        if (xx < 0 || xx >= W) continue;
        const float4 v = ld4(h1 + (((int64_t)b * H + yy) * W + xx) * C + c);
        const int t = ky * 3 + kx;
        const float4 wv = make_float4(w[(c + 0) * 9 + t], w[(c + 1) * 9 + t], w[(c + 2) * 9 + t], w[(c + 3) * 9 + t]);
        acc = fma4(v, wv, acc);
      }
    }
    const int64_t o = (((int64_t)b * H + y) * W + x) * C + c;
    st4(u2 + o, acc);
    if (h2) st4(h2 + o, make_float4(gelu_f(acc.x), gelu_f(acc.y), gelu_f(acc.z), gelu_f(acc.w)));
  }
}

// du1 = gelu'(u1) * sum_taps w[c][tap] * du2[p - delta(tap)]
__global__ void __launch_bounds__(256) dwconv_bwd_data_kernel(const float* __restrict__ du2,
                                                              const float* __restrict__ u1,
                                                              const float* __restrict__ w, float* __restrict__ du1,
                                                              int B, int H, int W, int C) {
  const int C4 = C >> 2;
  const int64_t total = (int64_t)B * H * W * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    int64_t p = i / C4;
    const int x = (int)(p % W); p /= W;
    const int y = (int)(p % H);
    const int b = (int)(p / H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y - (ky - 1);              // output pixel that read us through tap (ky,kx)
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x - (kx - 1);
        if (xx < 0 || xx >= W) continue;
        const float4 g = ld4(du2 + (((int64_t)b * H + yy) * W + xx) * C + c);
        const int t = ky * 3 + kx;
        const float4 wv = make_float4(w[(c + 0) * 9 + t], w[(c + 1) * 9 + t], w[(c + 2) * 9 + t], w[(c + 3) * 9 + t]);
        acc = fma4(g, wv, acc);
      }
    }
    const int64_t o = (((int64_t)b * H + y) * W + x) * C + c;
    if (u1) {
      const float4 u = ld4(u1 + o);
      acc = make_float4(acc.x * gelu_grad_f(u.x), acc.y * gelu_grad_f(u.y), acc.z * gelu_grad_f(u.z),
                        acc.w * gelu_grad_f(u.w));
    }
    st4(du1 + o, acc);
  }
}

// dw[c][tap] += sum_p du2[p,c] * h1[p + delta(tap), c];  db[c] += sum_p du2[p,c]
// block = 32 channel-quads x 8 pixel lanes, PIX pixels per block.
constexpr int DW_PIX = 512;
__global__ void __launch_bounds__(256) dwconv_bwd_weight_kernel(const float* __restrict__ du2,
                                                                const float* __restrict__ h1, float* __restrict__ dw,
                                                                float* __restrict__ db, int B, int H, int W, int C) {
  __shared__ float red[8][32][41];
  const int cq = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = (blockIdx.y * 32 + cq) * 4;
  const bool cok = c < C;
  const int64_t T = (int64_t)B * H * W;
  const int64_t p0 = (int64_t)blockIdx.x * DW_PIX;
  float acc[9][4];
  float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < 9; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
  if (cok) {
    for (int64_t p = p0 + pl; p < min(T, p0 + DW_PIX); p += 8) {
      const int x = (int)(p % W);
      const int y = (int)((p / W) % H);
      const float4 g = ld4(du2 + p * C + c);
      accb[0] += g.x; accb[1] += g.y; accb[2] += g.z; accb[3] += g.w;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = x + kx - 1;
          if (xx < 0 || xx >= W) continue;
          const float4 v = ld4(h1 + (p + (int64_t)(ky - 1) * W + (kx - 1)) * C + c);
          const int t = ky * 3 + kx;
          acc[t][0] = fmaf(g.x, v.x, acc[t][0]); acc[t][1] = fmaf(g.y, v.y, acc[t][1]);
          acc[t][2] = fmaf(g.z, v.z, acc[t][2]); acc[t][3] = fmaf(g.w, v.w, acc[t][3]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[pl][cq][t * 4 + j] = acc[t][j];
#pragma unroll
  for (int j = 0; j < 4; ++j) red[pl][cq][36 + j] = accb[j];
  __syncthreads();
  // 32 quads x 40 values reduced over the 8 pixel lanes
  for (int i = threadIdx.x; i < 32 * 40; i += 256) {
    const int q = i / 40, v = i % 40;
    const int cc = (blockIdx.y * 32 + q) * 4;
    if (cc >= C) continue;
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) s += red[l][q][v];
    if (v < 36) atomicAdd(&dw[(cc + (v & 3)) * 9 + (v >> 2)], s);
    else if (db) atomicAdd(&db[cc + (v - 36)], s);
  }
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

extern "C" {

int fa_dwconv3x3_fwd(const float* h1, const float* w, const float* b, float* u2, float* h2, int B, int H, int W, int C,
                     fa_stream_t stream) {
  FA_REQUIRE(h1 && w && u2, "fa_dwconv3x3_fwd: null pointer");
  FA_REQUIRE(C % 4 == 0, "fa_dwconv3x3_fwd: C=%d must be a multiple of 4", C);
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_DWCONV, st);
  const int64_t total = (int64_t)B * H * W * (C / 4);
  if (total == 0) return FA_OK;
  dwconv_fwd_kernel<<<ew_grid(total), 256, 0, st>>>(h1, w, b, u2, h2, B, H, W, C);
  FA_LAUNCH_CHECK("fa_dwconv3x3_fwd");
  return FA_OK;
}

int fa_dwconv3x3_bwd(const float* du2, const float* h1, const float* u1, const float* w, float* du1, float* dw,
                     float* db, int B, int H, int W, int C, fa_stream_t stream) {
  FA_REQUIRE(du2 && w && du1, "fa_dwconv3x3_bwd: null pointer");
  FA_REQUIRE(C % 4 == 0, "fa_dwconv3x3_bwd: C=%d must be a multiple of 4", C);
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_DWCONV, st);
  const int64_t total = (int64_t)B * H * W * (C / 4);
  if (total == 0) return FA_OK;
  dwconv_bwd_data_kernel<<<ew_grid(total), 256, 0, st>>>(du2, u1, w, du1, B, H, W, C);
  FA_LAUNCH_CHECK("fa_dwconv3x3_bwd(data)");
  if (dw) {
    FA_REQUIRE(h1, "fa_dwconv3x3_bwd: h1 required for the weight gradient");
    fa_count_launch(FA_K_DWCONV);
    const int64_t T = (int64_t)B * H * W;
    dim3 grid((unsigned)((T + DW_PIX - 1) / DW_PIX), (C / 4 + 31) / 32);
    dwconv_bwd_weight_kernel<<<grid, 256, 0, st>>>(du2, h1, dw, db, B, H, W, C);
    FA_LAUNCH_CHECK("fa_dwconv3x3_bwd(weight)");
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
