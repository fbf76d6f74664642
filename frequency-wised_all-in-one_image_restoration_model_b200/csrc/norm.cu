// LayerNorm (tokens) and BatchNorm2d(+LeakyReLU+avgpool) on [B,C,S]; column sums for bias gradients.
// HBM-bound kernels: one pass over the activations per call, 128-bit accesses where the row length allows,
// grids sized in multiples of the SM count.
#include "freqair_internal.h"

namespace {

// ---------------------------------------------------------------- LayerNorm
// One warp per row, lane-strided columns (MAXJ = ceil(C / 32) values per lane).  Narrow rows (C <= 128: the 128^2 and
// 64^2 levels, up to 786 432 rows of 112 bytes) are latency-bound at one row per warp iteration, so a warp walks
// ROWS rows at a time: the loads of all of them are issued before the first reduction.
template <int MAXJ, int ROWS>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd, int64_t rows,
                                                     int C) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float invC = 1.0f / (float)C;
  float gm[MAXJ], bt[MAXJ];
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    const int c = j * 32 + lane;
    gm[j] = (gamma && c < C) ? gamma[c] : 1.0f;
    bt[j] = (gamma && beta && c < C) ? beta[c] : 0.f;
  }
  for (int64_t r0 = warp * ROWS; r0 < rows; r0 += nwarps * ROWS) {
    float v[ROWS][MAXJ], s[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      const float* xr = x + (r0 + i) * C;
      const bool rok = r0 + i < rows;
      s[i] = 0.f;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int c = j * 32 + lane;
        v[i][j] = (rok && c < C) ? xr[c] : 0.f;
        s[i] += v[i][j];
      }
    }
    float mu[ROWS], q[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) mu[i] = warp_sum(s[i]) * invC;
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      q[i] = 0.f;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int c = j * 32 + lane;
        const float d = (c < C) ? v[i][j] - mu[i] : 0.f;
        q[i] += d * d;
      }
    }
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      if (r0 + i >= rows) break;
      const float rs = rsqrtf(warp_sum(q[i]) * invC + 1e-5f);
      float* yr = y + (r0 + i) * C;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int c = j * 32 + lane;
        if (c < C) yr[c] = fmaf((v[i][j] - mu[i]) * rs, gm[j], bt[j]);
      }
      if (lane == 0) {
        if (mean) mean[r0 + i] = mu[i];
        if (rstd) rstd[r0 + i] = rs;
      }
    }
  }
}

template <int MAXJ, int ROWS>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, const float* __restrict__ dres,
                                                     float* __restrict__ dx, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, int64_t rows, int C) {
  extern __shared__ float sm[];           // [2][C] block partials of dgamma / dbeta
  float* sg = sm;
  float* sb = sm + C;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float invC = 1.0f / (float)C;
  float ag[MAXJ], ab[MAXJ], gm[MAXJ];
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    ag[j] = 0.f; ab[j] = 0.f;
    int c = j * 32 + lane;
    gm[j] = (gamma && c < C) ? gamma[c] : 1.0f;
  }
  for (int64_t r0 = warp * ROWS; r0 < rows; r0 += nwarps * ROWS) {
    float g[ROWS][MAXJ], xh[ROWS][MAXJ], rsd[ROWS][MAXJ], s1[ROWS], s2[ROWS], rs[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      const bool rok = r0 + i < rows;
      const int64_t r = rok ? r0 + i : 0;
      const float mu = mean[r];
      rs[i] = rstd[r];
      const float* xr = x + r * C;
      const float* dr = dy + r * C;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int c = j * 32 + lane;
        const bool ok = rok && c < C;
        g[i][j] = ok ? dr[c] : 0.f;
        xh[i][j] = ok ? (xr[c] - mu) * rs[i] : 0.f;
        rsd[i][j] = (ok && dres) ? dres[r * C + c] : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      s1[i] = 0.f; s2[i] = 0.f;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const float d = g[i][j];
        ag[j] += d * xh[i][j];
        ab[j] += d;
        g[i][j] = d * gm[j];
        s1[i] += g[i][j];
        s2[i] += g[i][j] * xh[i][j];
      }
    }
#pragma unroll
    for (int i = 0; i < ROWS; ++i) { s1[i] = warp_sum(s1[i]) * invC; s2[i] = warp_sum(s2[i]) * invC; }
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      if (r0 + i >= rows) break;
      float* dxr = dx + (r0 + i) * C;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int c = j * 32 + lane;
        if (c < C) dxr[c] = fmaf(rs[i], g[i][j] - s1[i] - xh[i][j] * s2[i], rsd[i][j]);
      }
    }
  }
  if (dgamma || dbeta) {
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      int c = j * 32 + lane;
      if (c < C) { atomicAdd(&sg[c], ag[j]); atomicAdd(&sb[c], ab[j]); }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dgamma) atomicAdd(&dgamma[c], sg[c]);
      if (dbeta) atomicAdd(&dbeta[c], sb[c]);
    }
  }
}

// ---------------------------------------------------------------- column sums
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, float* __restrict__ out, int M, int N,
                                                     int64_t ld, const float* __restrict__ rowscale, int rps,
                                                     int rows_per_block) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int m0 = blockIdx.y * rows_per_block;
  const int m1 = min(M, m0 + rows_per_block);
  float s = 0.f;
  if (n < N) {
    for (int m = m0 + ty; m < m1; m += 8) {
      float v = X[(int64_t)m * ld + n];
      if (rowscale) v *= rowscale[m / rps];
      s += v;
    }
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    atomicAdd(&out[n], t);
  }
}

// out[m] += sum_k X[m*ld + k]  (one warp per row)
__global__ void __launch_bounds__(256) rowsum_kernel(const float* __restrict__ X, float* __restrict__ out, int M, int K,
                                                     int64_t ld) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t m = warp; m < M; m += nwarps) {
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s += X[m * ld + k];
    s = warp_sum(s);
    if (lane == 0) out[m] += s;
  }
}

// ---------------------------------------------------------------- BatchNorm on [B,C,S]
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
  if (w == 0) t = warp_sum(t);
  __syncthreads();
  return t;       // valid in warp 0
}

constexpr int BN_CHUNK = 4096;

__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, double* __restrict__ sums, int C,
                                                       int64_t S, int chunks) {
  __shared__ float sh[8];
  const int64_t bc = blockIdx.x / chunks;
  const int ch = blockIdx.x % chunks;
  const int c = (int)(bc % C);
  const int64_t s0 = (int64_t)ch * BN_CHUNK, s1 = min(S, s0 + BN_CHUNK);
  const float* p = x + bc * S;
  float a = 0.f, q = 0.f;
  for (int64_t s = s0 + threadIdx.x; s < s1; s += blockDim.x) { float v = p[s]; a += v; q += v * v; }
  a = block_sum(a, sh);
  q = block_sum(q, sh);
  if (threadIdx.x == 0) { atomicAdd(&sums[2 * c], (double)a); atomicAdd(&sums[2 * c + 1], (double)q); }
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, float slope,
                                                       float* __restrict__ y, float* __restrict__ pooled, int C,
                                                       int64_t S) {
  __shared__ float sh[8];
  const int64_t bc = blockIdx.x;
  const int c = (int)(bc % C);
  const float sc = scale[c], sf = shift[c];
  const float* p = x + bc * S;
  float acc = 0.f;
  for (int64_t s = threadIdx.x; s < S; s += blockDim.x) {
    float z = p[s] * sc + sf;
    z = z > 0.f ? z : z * slope;
    if (y) y[bc * S + s] = z;
    acc += z;
  }
  if (pooled) {
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) pooled[bc] = acc / (float)S;
  }
}

// g = upstream * lrelu'(z); upstream = dy[b,c,s] if dy else dpooled[b,c]/S
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift, float slope,
                                                            const float* __restrict__ dy,
                                                            const float* __restrict__ dpooled, double* __restrict__ red,
                                                            int C, int64_t S, int chunks) {
  __shared__ float sh[8];
  const int64_t bc = blockIdx.x / chunks;
  const int ch = blockIdx.x % chunks;
  const int c = (int)(bc % C);
  const int64_t s0 = (int64_t)ch * BN_CHUNK, s1 = min(S, s0 + BN_CHUNK);
  const float sc = scale[c], sf = shift[c], mu = mean[c], rs = rstd[c];
  const float up = dpooled ? dpooled[bc] / (float)S : 0.f;
  const float* p = x + bc * S;
  float a = 0.f, q = 0.f;
  for (int64_t s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
    float v = p[s];
    float z = v * sc + sf;
    float g = (dy ? dy[bc * S + s] : up) * (z > 0.f ? 1.0f : slope);
    a += g;
    q += g * (v - mu) * rs;
  }
  a = block_sum(a, sh);
  q = block_sum(q, sh);
  if (threadIdx.x == 0) { atomicAdd(&red[2 * c], (double)a); atomicAdd(&red[2 * c + 1], (double)q); }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift, float slope,
                                                           const float* __restrict__ dy,
                                                           const float* __restrict__ dpooled,
                                                           const double* __restrict__ red, float* __restrict__ dx,
                                                           int B, int C, int64_t S) {
  const int64_t bc = blockIdx.x;
  const int c = (int)(bc % C);
  const float sc = scale[c], sf = shift[c], mu = mean[c], rs = rstd[c];
  const float up = dpooled ? dpooled[bc] / (float)S : 0.f;
  const double n = (double)B * (double)S;
  const float r0 = red ? (float)(red[2 * c] / n) : 0.f;
  const float r1 = red ? (float)(red[2 * c + 1] / n) : 0.f;
  const float* p = x + bc * S;
  for (int64_t s = threadIdx.x; s < S; s += blockDim.x) {
    float v = p[s];
    float z = v * sc + sf;
    float g = (dy ? dy[bc * S + s] : up) * (z > 0.f ? 1.0f : slope);
    float xh = (v - mu) * rs;
    dx[bc * S + s] = sc * (g - r0 - xh * r1);
  }
}

// ---------------------------------------------------------------- BatchNorm on tokens [T, C] (NHWC)
// block = 32 channels x 8 row lanes; rows_per_block rows per block; double atomics on the [C][2] result.
__global__ void __launch_bounds__(256) bnt_stats_kernel(const float* __restrict__ x, double* __restrict__ sums,
                                                        int64_t T, int C, int rows_per_block) {
  __shared__ float ra[8][33], rq[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(T, r0 + rows_per_block);
  float a = 0.f, q = 0.f;
  if (c < C)
    for (int64_t r = r0 + ty; r < r1; r += 8) { const float v = x[r * C + c]; a += v; q += v * v; }
  ra[ty][tx] = a; rq[ty][tx] = q;
  __syncthreads();
  if (ty == 0 && c < C) {
    float sa = 0.f, sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { sa += ra[i][tx]; sq += rq[i][tx]; }
    atomicAdd(&sums[2 * c], (double)sa);
    atomicAdd(&sums[2 * c + 1], (double)sq);
  }
}

// y = act(x*scale[c] + shift[c] + res)
__global__ void __launch_bounds__(256) bnt_apply_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                        const float* __restrict__ shift,
                                                        const float* __restrict__ res, float slope,
                                                        float* __restrict__ y, int64_t total, int C) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float z = x[i] * scale[c] + shift[c];
    if (res) z += res[i];
    y[i] = z > 0.f ? z : z * slope;
  }
}

// g = dy * act'(y);  red[c] += {sum g, sum g*xhat}
__global__ void __launch_bounds__(256) bnt_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd,
                                                             const float* __restrict__ yout, float slope,
                                                             const float* __restrict__ dy, double* __restrict__ red,
                                                             int64_t T, int C, int rows_per_block) {
  __shared__ float ra[8][33], rq[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(T, r0 + rows_per_block);
  float a = 0.f, q = 0.f;
  if (c < C) {
    const float mu = mean[c], rs = rstd[c];
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const int64_t i = r * C + c;
      float g = dy[i];
      if (yout && !(yout[i] > 0.f)) g *= slope;
      a += g;
      q += g * (x[i] - mu) * rs;
    }
  }
  ra[ty][tx] = a; rq[ty][tx] = q;
  __syncthreads();
  if (ty == 0 && c < C) {
    float sa = 0.f, sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { sa += ra[i][tx]; sq += rq[i][tx]; }
    atomicAdd(&red[2 * c], (double)sa);
    atomicAdd(&red[2 * c + 1], (double)sq);
  }
}

__global__ void __launch_bounds__(256) bnt_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ yout, float slope,
                                                            const float* __restrict__ dy, const double* __restrict__ red,
                                                            float* __restrict__ dx, float* __restrict__ dres,
                                                            int64_t T, int C) {
  const int64_t total = T * C;
  const double n = (double)T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float g = dy[i];
    if (yout && !(yout[i] > 0.f)) g *= slope;
    if (dres) dres[i] = g;
    const float xh = (x[i] - mean[c]) * rstd[c];
    const float r0 = red ? (float)(red[2 * c] / n) : 0.f, r1 = red ? (float)(red[2 * c + 1] / n) : 0.f;
    dx[i] = scale[c] * (g - r0 - xh * r1);
  }
}

// out[b][c] = mean_t x[b][t][c]   /   dx[b][t][c] = dy[b][c] / HW
__global__ void __launch_bounds__(256) token_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int HW,
                                                         int C) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int b = blockIdx.y;
  float a = 0.f;
  if (c < C)
    for (int t = ty; t < HW; t += 8) a += x[((int64_t)b * HW + t) * C + c];
  red[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][tx];
    out[(int64_t)b * C + c] = s / (float)HW;
  }
}
__global__ void __launch_bounds__(256) token_mean_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                                             int64_t total, int HW, int C) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t b = i / ((int64_t)HW * C);
    dx[i] = dy[b * C + c] / (float)HW;
  }
}

int ln_grid(int64_t rows) {
  int64_t blocks = (rows + 7) / 8;
  int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

extern "C" {

int fa_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                     int64_t rows, int C, fa_stream_t stream) {
  FA_REQUIRE(x && y, "fa_layernorm_fwd: null pointer");
  FA_REQUIRE(C >= 1 && C <= 1024, "fa_layernorm_fwd: C=%d unsupported (1..1024)", C);
  if (rows == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_LAYERNORM, st);
  const int grid = ln_grid(rows);
  if (C <= 32) ln_fwd_kernel<1, 8><<<ln_grid(rows / 8 + 1), 256, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, C);
  else if (C <= 64) ln_fwd_kernel<2, 4><<<ln_grid(rows / 4 + 1), 256, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, C);
  else if (C <= 128) ln_fwd_kernel<4, 4><<<ln_grid(rows / 4 + 1), 256, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, C);
  else if (C <= 256) ln_fwd_kernel<8, 2><<<ln_grid(rows / 2 + 1), 256, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, C);
  else if (C <= 512) ln_fwd_kernel<16, 1><<<grid, 256, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, C);
  else ln_fwd_kernel<32, 1><<<grid, 256, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, C);
  FA_LAUNCH_CHECK("fa_layernorm_fwd");
  return FA_OK;
}

int fa_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                     const float* dres, float* dx, float* dgamma, float* dbeta, int64_t rows, int C,
                     fa_stream_t stream) {
  FA_REQUIRE(dy && x && mean && rstd && dx, "fa_layernorm_bwd: null pointer");
  FA_REQUIRE(C >= 1 && C <= 1024, "fa_layernorm_bwd: C=%d unsupported (1..1024)", C);
  if (rows == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_LAYERNORM, st);
  const size_t smem = 2 * (size_t)C * sizeof(float);
  auto grid_for = [&](int rows_per_warp) {
    int g = ln_grid(rows / rows_per_warp + 1);
    return g > 4 * kNumSMs ? 4 * kNumSMs : g;
  };
  // rows per warp iteration measured on B200 (786432x28: 1 -> 0.182 ms, 2 -> 0.135, 4 -> 0.127, 8 -> 0.258;
  // 262144x56: 1 -> 0.064, 2 -> 0.066, 4 -> 0.115)
  if (C <= 32) ln_bwd_kernel<1, 4><<<grid_for(4), 256, smem, st>>>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, rows, C);
  else if (C <= 64) ln_bwd_kernel<2, 1><<<grid_for(1), 256, smem, st>>>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, rows, C);
  else if (C <= 128) ln_bwd_kernel<4, 2><<<grid_for(2), 256, smem, st>>>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, rows, C);
  else if (C <= 256) ln_bwd_kernel<8, 1><<<grid_for(1), 256, smem, st>>>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, rows, C);
  else if (C <= 512) ln_bwd_kernel<16, 1><<<grid_for(1), 256, smem, st>>>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, rows, C);
  else ln_bwd_kernel<32, 1><<<grid_for(1), 256, smem, st>>>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, rows, C);
  FA_LAUNCH_CHECK("fa_layernorm_bwd");
  return FA_OK;
}

int fa_colsum(const float* X, float* out, int M, int N, int64_t ld, const float* rowscale, int rows_per_scale,
              int accumulate, fa_stream_t stream) {
  FA_REQUIRE(X && out, "fa_colsum: null pointer");
  if (N <= 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  if (!accumulate) FA_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N, st));
  if (M <= 0) return FA_OK;
  const int colblocks = (N + 31) / 32;
  int rowblocks = (4 * kNumSMs + colblocks - 1) / colblocks;
  int rpb = (M + rowblocks - 1) / rowblocks;
  if (rpb < 64) rpb = 64;
  rowblocks = (M + rpb - 1) / rpb;
  colsum_kernel<<<dim3(colblocks, rowblocks), 256, 0, st>>>(X, out, M, N, ld, rowscale,
                                                            rows_per_scale > 0 ? rows_per_scale : 1, rpb);
  FA_LAUNCH_CHECK("fa_colsum");
  return FA_OK;
}

int fa_a_rowsum_impl(const float* A, float* out, int M, int K, int64_t lda, int transA, fa_stream_t stream) {
  if (transA) return fa_colsum(A, out, K, M, lda, nullptr, 1, 1, stream);     // stored [K, M]: column sums
  cudaStream_t st = (cudaStream_t)stream;
  if (M <= 0 || K <= 0) return FA_OK;
  const int grid = ln_grid(M);
  rowsum_kernel<<<grid, 256, 0, st>>>(A, out, M, K, lda);
  FA_LAUNCH_CHECK("fa_gemm(a_rowsum)");
  return FA_OK;
}

int fa_bn_stats(const float* x, double* sums, int B, int C, int64_t S, fa_stream_t stream) {
  FA_REQUIRE(x && sums && B > 0 && C > 0 && S > 0, "fa_bn_stats: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_BN, st);
  const int chunks = (int)((S + BN_CHUNK - 1) / BN_CHUNK);
  bn_stats_kernel<<<(unsigned)((int64_t)B * C * chunks), 256, 0, st>>>(x, sums, C, S, chunks);
  FA_LAUNCH_CHECK("fa_bn_stats");
  return FA_OK;
}

int fa_bn_apply(const float* x, const float* scale, const float* shift, float slope, float* y, float* pooled, int B,
                int C, int64_t S, fa_stream_t stream) {
  FA_REQUIRE(x && scale && shift && (y || pooled) && B > 0 && C > 0 && S > 0, "fa_bn_apply: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_BN, st);
  bn_apply_kernel<<<(unsigned)((int64_t)B * C), 256, 0, st>>>(x, scale, shift, slope, y, pooled, C, S);
  FA_LAUNCH_CHECK("fa_bn_apply");
  return FA_OK;
}

int fa_bn_bwd_reduce(const float* x, const float* mean, const float* rstd, const float* scale, const float* shift,
                     float slope, const float* dy, const float* dpooled, double* red, int B, int C, int64_t S,
                     fa_stream_t stream) {
  FA_REQUIRE(x && mean && rstd && scale && shift && red && (dy || dpooled), "fa_bn_bwd_reduce: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_BN, st);
  const int chunks = (int)((S + BN_CHUNK - 1) / BN_CHUNK);
  bn_bwd_reduce_kernel<<<(unsigned)((int64_t)B * C * chunks), 256, 0, st>>>(
      x, mean, rstd, scale, shift, slope, dy, dpooled, red, C, S, chunks);
  FA_LAUNCH_CHECK("fa_bn_bwd_reduce");
  return FA_OK;
}

int fa_bn_bwd_apply(const float* x, const float* mean, const float* rstd, const float* scale, const float* shift,
                    float slope, const float* dy, const float* dpooled, const double* red, float* dx, int B, int C,
                    int64_t S, fa_stream_t stream) {
  FA_REQUIRE(x && mean && rstd && scale && shift && dx && (dy || dpooled), "fa_bn_bwd_apply: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_BN, st);
  bn_bwd_apply_kernel<<<(unsigned)((int64_t)B * C), 256, 0, st>>>(x, mean, rstd, scale, shift, slope, dy, dpooled,
                                                                  red, dx, B, C, S);
  FA_LAUNCH_CHECK("fa_bn_bwd_apply");
  return FA_OK;
}

static void bnt_grid(int64_t T, int C, dim3& grid, int& rpb) {
  const int cb = (C + 31) / 32;
  int rb = (4 * kNumSMs + cb - 1) / cb;
  rpb = (int)((T + rb - 1) / rb);
  if (rpb < 64) rpb = 64;
  rb = (int)((T + rpb - 1) / rpb);
  grid = dim3(cb, rb);
}
static int ew_blocks(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b > cap ? cap : (b > 0 ? b : 1));
}

int fa_bn_tokens_stats(const float* x, double* sums, int64_t T, int C, fa_stream_t stream) {
  FA_REQUIRE(x && sums && T > 0 && C > 0, "fa_bn_tokens_stats: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_BN, st);
  dim3 grid; int rpb;
  bnt_grid(T, C, grid, rpb);
  bnt_stats_kernel<<<grid, 256, 0, st>>>(x, sums, T, C, rpb);
  FA_LAUNCH_CHECK("fa_bn_tokens_stats");
  return FA_OK;
}

int fa_bn_tokens_apply(const float* x, const float* scale, const float* shift, const float* res, float slope, float* y,
                       int64_t T, int C, fa_stream_t stream) {
  FA_REQUIRE(x && scale && shift && y && T > 0 && C > 0, "fa_bn_tokens_apply: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_BN, st);
  bnt_apply_kernel<<<ew_blocks(T * C), 256, 0, st>>>(x, scale, shift, res, slope, y, T * C, C);
  FA_LAUNCH_CHECK("fa_bn_tokens_apply");
  return FA_OK;
}

int fa_bn_tokens_bwd(const float* x, const float* mean, const float* rstd, const float* scale, const float* yout,
                     float slope, const float* dy, double* red, float* dx, float* dres, int64_t T, int C, int training,
                     fa_stream_t stream) {
  FA_REQUIRE(x && mean && rstd && scale && dy && red && dx && T > 0 && C > 0, "fa_bn_tokens_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_BN, st);
  dim3 grid; int rpb;
  bnt_grid(T, C, grid, rpb);
  bnt_bwd_reduce_kernel<<<grid, 256, 0, st>>>(x, mean, rstd, yout, slope, dy, red, T, C, rpb);
  FA_LAUNCH_CHECK("fa_bn_tokens_bwd(reduce)");
  fa_count_launch(FA_K_BN);
  bnt_bwd_apply_kernel<<<ew_blocks(T * C), 256, 0, st>>>(x, mean, rstd, scale, yout, slope, dy, training ? red : nullptr,
                                                        dx, dres, T, C);
  FA_LAUNCH_CHECK("fa_bn_tokens_bwd(apply)");
  return FA_OK;
}

int fa_token_mean_fwd(const float* x, float* out, int B, int HW, int C, fa_stream_t stream) {
  FA_REQUIRE(x && out && B > 0 && HW > 0 && C > 0, "fa_token_mean_fwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  token_mean_kernel<<<dim3((C + 31) / 32, B), 256, 0, st>>>(x, out, HW, C);
  FA_LAUNCH_CHECK("fa_token_mean_fwd");
  return FA_OK;
}

int fa_token_mean_bwd(const float* dy, float* dx, int B, int HW, int C, fa_stream_t stream) {
  FA_REQUIRE(dy && dx && B > 0 && HW > 0 && C > 0, "fa_token_mean_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  const int64_t total = (int64_t)B * HW * C;
  token_mean_bwd_kernel<<<ew_blocks(total), 256, 0, st>>>(dy, dx, total, HW, C);
  FA_LAUNCH_CHECK("fa_token_mean_bwd");
  return FA_OK;
}

}  // extern "C"

int fa_a_rowsum(const float* A, float* out, int M, int K, int64_t lda, int transA, fa_stream_t stream) {
  return fa_a_rowsum_impl(A, out, M, K, lda, transA, stream);
}
