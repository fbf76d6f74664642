// Pure-bandwidth kernels: activations, L1 loss, MoCo momentum update, Adam, SFT/DGM fusion.
// All stream a flat buffer once with 128-bit accesses; grid = multiple of the SM count, grid-stride loops.
#include "freqair_internal.h"

namespace {

inline int ew_grid(int64_t n4) {
  int64_t blocks = (n4 + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  if (blocks > cap) blocks = cap;
  return (int)(blocks > 0 ? blocks : 1);
}

#define FA_GRID_STRIDE(i, n) \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

__global__ void __launch_bounds__(256) act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n,
                                                      int act, float p) {
  FA_GRID_STRIDE(i, n) y[i] = act_f(x[i], act, p);
}
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                      float* __restrict__ dx, int64_t n, int act, float p) {
  FA_GRID_STRIDE(i, n) dx[i] = dy[i] * act_grad_f(x[i], act, p);
}

__global__ void __launch_bounds__(256) l1_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                 float* __restrict__ loss, float* __restrict__ grad, int64_t n,
                                                 float gscale) {
  __shared__ float sh[8];
  float acc = 0.f;
  const float inv = 1.0f / (float)n;
  FA_GRID_STRIDE(i, n) {
    const float d = a[i] - b[i];
    acc += fabsf(d);
    if (grad) grad[i] = (d > 0.f ? 1.0f : (d < 0.f ? -1.0f : 0.0f)) * gscale * inv;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i];
    atomicAdd(loss, t * inv);
  }
}

__global__ void __launch_bounds__(256) momentum_kernel(float4* __restrict__ k, const float4* __restrict__ q, int64_t n4,
                                                       float m) {
  const float om = 1.0f - m;
  FA_GRID_STRIDE(i, n4) {
    float4 a = k[i];
    const float4 b = q[i];
    a.x = a.x * m + b.x * om; a.y = a.y * m + b.y * om; a.z = a.z * m + b.z * om; a.w = a.w * m + b.w * om;
    k[i] = a;
  }
}
__global__ void __launch_bounds__(256) momentum_tail_kernel(float* __restrict__ k, const float* __restrict__ q,
                                                            int64_t beg, int64_t n, float m) {
  FA_GRID_STRIDE(j, n - beg) { const int64_t i = beg + j; k[i] = k[i] * m + q[i] * (1.0f - m); }
}

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float b1, float b2, float eps,
                                      float step_size, float inv_bc2_sqrt) {
  m = m * b1 + g * (1.0f - b1);
  v = v * b2 + g * g * (1.0f - b2);
  const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n, float b1,
                                                   float b2, float eps, float step_size, float inv_bc2_sqrt,
                                                   float gscale, int vec, const float* __restrict__ state) {
  if (state) {
    // {lr, step count} live in device memory (fa_adam_tick advances the count): the bias corrections of
    // torch.optim.Adam are derived here, in double like torch's host code, once per block
    __shared__ float sh[2];
    if (threadIdx.x == 0) {
      const double t = (double)reinterpret_cast<const int*>(state)[1];
      const double bc1 = 1.0 - pow((double)b1, t), bc2 = 1.0 - pow((double)b2, t);
      sh[0] = (float)((double)state[0] / bc1);
      sh[1] = (float)(1.0 / sqrt(bc2));
    }
    __syncthreads();
    step_size = sh[0]; inv_bc2_sqrt = sh[1];
  }
  if (vec) {
    const int64_t n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    FA_GRID_STRIDE(i, n4) {
      float4 pp = p4[i], mm = m4[i], vv = v4[i];
      const float4 gg = g4[i];
      adam1(pp.x, gg.x * gscale, mm.x, vv.x, b1, b2, eps, step_size, inv_bc2_sqrt);
      adam1(pp.y, gg.y * gscale, mm.y, vv.y, b1, b2, eps, step_size, inv_bc2_sqrt);
      adam1(pp.z, gg.z * gscale, mm.z, vv.z, b1, b2, eps, step_size, inv_bc2_sqrt);
      adam1(pp.w, gg.w * gscale, mm.w, vv.w, b1, b2, eps, step_size, inv_bc2_sqrt);
      p4[i] = pp; m4[i] = mm; v4[i] = vv;
    }
    FA_GRID_STRIDE(j, n - (n4 << 2)) {
      const int64_t i = (n4 << 2) + j;
      adam1(p[i], g[i] * gscale, m[i], v[i], b1, b2, eps, step_size, inv_bc2_sqrt);
    }
  } else {
    FA_GRID_STRIDE(i, n) adam1(p[i], g[i] * gscale, m[i], v[i], b1, b2, eps, step_size, inv_bc2_sqrt);
  }
}

// out = act(x + dcn + x*gamma + beta)
__global__ void __launch_bounds__(256) sft_fwd_kernel(const float* __restrict__ x, const float* __restrict__ dcn,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      float* __restrict__ out, int64_t n, float slope) {
  FA_GRID_STRIDE(i, n) {
    const float z = x[i] + dcn[i] + x[i] * gamma[i] + beta[i];
    out[i] = z > 0.f ? z : z * slope;
  }
}
__global__ void __launch_bounds__(256) sft_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dcn,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ dout, float* __restrict__ dx,
                                                      float* __restrict__ ddcn, float* __restrict__ dgamma,
                                                      float* __restrict__ dbeta, int64_t n, float slope) {
  FA_GRID_STRIDE(i, n) {
    const float z = x[i] + dcn[i] + x[i] * gamma[i] + beta[i];
    const float g = dout[i] * (z > 0.f ? 1.0f : slope);
    dx[i] = g * (1.0f + gamma[i]);
    ddcn[i] = g;
    dgamma[i] = g * x[i];
    dbeta[i] = g;
  }
}

// ---------------------------------------------------------------- band-weight (lambda) predictor of one decoder block
// e = fc_w (s * ln_w + ln_b) + fc_b ;  h = lrelu(W0 e + b0, 0.1) ;  out = W2 h + b2      (decoder_Uformer.py:178-193,280-284
// with the LayerNorm statistics and the token mean hoisted out: s = mean_tokens(LN_noaffine(inter)), shared by all blocks)
// One CTA per sample; heads <= 32, D <= 1024.  ~10 torch kernels forward and ~25 backward per (block, band) otherwise.
struct CoefP {
  const float *ln_w, *ln_b, *fc_w, *fc_b, *w0, *b0, *w2, *b2;
};
struct CoefG {
  float *ln_w, *ln_b, *fc_w, *fc_b, *w0, *b0, *w2, *b2;
};

__device__ __forceinline__ void coef_forward(const float* __restrict__ s, const CoefP& p, int D, int heads, float* z,
                                             float* e, float* pre, float* h) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int d = tid; d < D; d += blockDim.x) z[d] = s[d] * p.ln_w[d] + p.ln_b[d];
  __syncthreads();
  for (int j = w; j < heads; j += (blockDim.x >> 5)) {
    float a = 0.f;
    for (int d = lane; d < D; d += 32) a = fmaf(p.fc_w[(int64_t)j * D + d], z[d], a);
    a = warp_sum(a);
    if (lane == 0) e[j] = a + p.fc_b[j];
  }
  __syncthreads();
  if (tid < heads) {
    float a = p.b0[tid];
    for (int i = 0; i < heads; ++i) a = fmaf(p.w0[tid * heads + i], e[i], a);
    pre[tid] = a;
    h[tid] = a > 0.f ? a : 0.1f * a;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) band_coef_fwd_kernel(const float* __restrict__ stats, CoefP p, float* __restrict__ out,
                                                            int D, int heads, int64_t ld_b, int64_t ld_h) {
  __shared__ float z[1024], e[32], pre[32], h[32];
  const int b = blockIdx.x;
  coef_forward(stats + (int64_t)b * D, p, D, heads, z, e, pre, h);
  if (threadIdx.x < heads) {
    float a = p.b2[threadIdx.x];
    for (int i = 0; i < heads; ++i) a = fmaf(p.w2[threadIdx.x * heads + i], h[i], a);
    out[b * ld_b + threadIdx.x * ld_h] = a;
  }
}

__global__ void __launch_bounds__(256) band_coef_bwd_kernel(const float* __restrict__ stats, CoefP p,
                                                            const float* __restrict__ dout, int64_t ld_b, int64_t ld_h,
                                                            float* __restrict__ dstats, CoefG g, int D, int heads) {
  __shared__ float z[1024], e[32], pre[32], h[32], dh[32], dpre[32], de[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* s = stats + (int64_t)b * D;
  coef_forward(s, p, D, heads, z, e, pre, h);
  if (tid < heads) {
    const float go = dout[b * ld_b + tid * ld_h];
    atomicAdd(&g.b2[tid], go);
    for (int i = 0; i < heads; ++i) atomicAdd(&g.w2[tid * heads + i], go * h[i]);
    dh[tid] = go;                 // reused below as dout
  }
  __syncthreads();
  if (tid < heads) {
    float a = 0.f;
    for (int j = 0; j < heads; ++j) a = fmaf(p.w2[j * heads + tid], dh[j], a);
    dpre[tid] = a * (pre[tid] > 0.f ? 1.0f : 0.1f);
  }
  __syncthreads();
  if (tid < heads) {
    atomicAdd(&g.b0[tid], dpre[tid]);
    for (int i = 0; i < heads; ++i) atomicAdd(&g.w0[tid * heads + i], dpre[tid] * e[i]);
    float a = 0.f;
    for (int j = 0; j < heads; ++j) a = fmaf(p.w0[j * heads + tid], dpre[j], a);
    de[tid] = a;
    atomicAdd(&g.fc_b[tid], a);
  }
  __syncthreads();
  for (int d = tid; d < D; d += blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < heads; ++j) {
      atomicAdd(&g.fc_w[(int64_t)j * D + d], de[j] * z[d]);
      a = fmaf(p.fc_w[(int64_t)j * D + d], de[j], a);
    }
    atomicAdd(&g.ln_w[d], a * s[d]);
    atomicAdd(&g.ln_b[d], a);
    if (dstats) atomicAdd(&dstats[(int64_t)b * D + d], a * p.ln_w[d]);
  }
}

}  // namespace

extern "C" {

int fa_band_coef_fwd(const float* stats, const float* const* params, float* out, int B, int D, int heads, int64_t ld_b,
                     int64_t ld_h, fa_stream_t stream) {
  FA_REQUIRE(stats && params && out, "fa_band_coef_fwd: null pointer");
  FA_REQUIRE(D >= 1 && D <= 1024 && heads >= 1 && heads <= 32, "fa_band_coef_fwd: D=%d (<=1024), heads=%d (<=32)", D, heads);
  for (int i = 0; i < 8; ++i) FA_REQUIRE(params[i], "fa_band_coef_fwd: null parameter %d", i);
  if (B == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  CoefP p{params[0], params[1], params[2], params[3], params[4], params[5], params[6], params[7]};
  band_coef_fwd_kernel<<<B, 256, 0, st>>>(stats, p, out, D, heads, ld_b, ld_h);
  FA_LAUNCH_CHECK("fa_band_coef_fwd");
  return FA_OK;
}

int fa_band_coef_bwd(const float* stats, const float* const* params, const float* dout, int64_t ld_b, int64_t ld_h,
                     float* dstats, float* const* grads, int B, int D, int heads, fa_stream_t stream) {
  FA_REQUIRE(stats && params && dout && grads, "fa_band_coef_bwd: null pointer");
  FA_REQUIRE(D >= 1 && D <= 1024 && heads >= 1 && heads <= 32, "fa_band_coef_bwd: D=%d (<=1024), heads=%d (<=32)", D, heads);
  for (int i = 0; i < 8; ++i) FA_REQUIRE(params[i] && grads[i], "fa_band_coef_bwd: null parameter / gradient %d", i);
  if (B == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  CoefP p{params[0], params[1], params[2], params[3], params[4], params[5], params[6], params[7]};
  CoefG g{grads[0], grads[1], grads[2], grads[3], grads[4], grads[5], grads[6], grads[7]};
  band_coef_bwd_kernel<<<B, 256, 0, st>>>(stats, p, dout, ld_b, ld_h, dstats, g, D, heads);
  FA_LAUNCH_CHECK("fa_band_coef_bwd");
  return FA_OK;
}


int fa_act_fwd(const float* x, float* y, int64_t n, int act, float p, fa_stream_t stream) {
  FA_REQUIRE(x && y, "fa_act_fwd: null pointer");
  if (n == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  act_fwd_kernel<<<ew_grid(n), 256, 0, st>>>(x, y, n, act, p);
  FA_LAUNCH_CHECK("fa_act_fwd");
  return FA_OK;
}

int fa_act_bwd(const float* dy, const float* x, float* dx, int64_t n, int act, float p, fa_stream_t stream) {
  FA_REQUIRE(dy && x && dx, "fa_act_bwd: null pointer");
  if (n == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  act_bwd_kernel<<<ew_grid(n), 256, 0, st>>>(dy, x, dx, n, act, p);
  FA_LAUNCH_CHECK("fa_act_bwd");
  return FA_OK;
}

int fa_l1_loss(const float* a, const float* b, float* loss, float* grad, int64_t n, float gscale, fa_stream_t stream) {
  FA_REQUIRE(a && b && loss && n > 0, "fa_l1_loss: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  int grid = ew_grid(n);
  if (grid > 2 * kNumSMs) grid = 2 * kNumSMs;
  l1_kernel<<<grid, 256, 0, st>>>(a, b, loss, grad, n, gscale);
  FA_LAUNCH_CHECK("fa_l1_loss");
  return FA_OK;
}

int fa_momentum_update(float* k, const float* q, int64_t n, float m, fa_stream_t stream) {
  FA_REQUIRE(k && q, "fa_momentum_update: null pointer");
  if (n == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_OPTIM, st);
  const bool al = (((uintptr_t)k | (uintptr_t)q) % 16) == 0;
  const int64_t n4 = al ? (n >> 2) : 0;
  if (n4) momentum_kernel<<<ew_grid(n4), 256, 0, st>>>(reinterpret_cast<float4*>(k), reinterpret_cast<const float4*>(q), n4, m);
  if ((n4 << 2) < n) momentum_tail_kernel<<<ew_grid(n - (n4 << 2)), 256, 0, st>>>(k, q, n4 << 2, n, m);
  FA_LAUNCH_CHECK("fa_momentum_update");
  return FA_OK;
}

int fa_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                 int step, float grad_scale, fa_stream_t stream) {
  FA_REQUIRE(p && g && m && v && step >= 1, "fa_adam_step: bad argument");
  if (n == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_OPTIM, st);
  // torch.optim.Adam: step_size = lr / (1 - b1^t); denom = sqrt(v) / sqrt(1 - b2^t) + eps
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  const bool al = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16) == 0;
  adam_kernel<<<ew_grid(n / 4 + 1), 256, 0, st>>>(p, g, m, v, n, beta1, beta2, eps, step_size, inv_bc2_sqrt, grad_scale,
                                                  al ? 1 : 0, nullptr);
  FA_LAUNCH_CHECK("fa_adam_step");
  return FA_OK;
}

// dst = x rounded to the nearest TF32 (ties away), low 13 mantissa bits cleared: exactly what a 1xTF32 / 2xTF32 contraction
// wants as its weight operand (FaGemmEpilogue::b_is_tf32)
__device__ __forceinline__ float round_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__global__ void __launch_bounds__(256) round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n,
                                                         int vec) {
  if (vec) {
    const int64_t n4 = n >> 2;
    FA_GRID_STRIDE(i, n4) {
      const float4 v = reinterpret_cast<const float4*>(src)[i];
      reinterpret_cast<float4*>(dst)[i] = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
    }
    FA_GRID_STRIDE(j, n - (n4 << 2)) dst[(n4 << 2) + j] = round_tf32(src[(n4 << 2) + j]);
  } else {
    FA_GRID_STRIDE(i, n) dst[i] = round_tf32(src[i]);
  }
}

__global__ void adam_tick_kernel(float* state) { reinterpret_cast<int*>(state)[1] += 1; }

int fa_round_tf32(const float* src, float* dst, int64_t n, fa_stream_t stream) {
  FA_REQUIRE(src && dst && n >= 0, "fa_round_tf32: bad argument");
  if (n == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  const bool al = (((uintptr_t)src | (uintptr_t)dst) % 16) == 0;
  round_tf32_kernel<<<ew_grid(n / 4 + 1), 256, 0, st>>>(src, dst, n, al ? 1 : 0);
  FA_LAUNCH_CHECK("fa_round_tf32");
  return FA_OK;
}

int fa_adam_tick(float* state, fa_stream_t stream) {
  FA_REQUIRE(state, "fa_adam_tick: null state");
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_OPTIM, st);
  adam_tick_kernel<<<1, 1, 0, st>>>(state);
  FA_LAUNCH_CHECK("fa_adam_tick");
  return FA_OK;
}

int fa_adam_step_state(float* p, const float* g, float* m, float* v, int64_t n, const float* state, float beta1,
                       float beta2, float eps, float grad_scale, fa_stream_t stream) {
  FA_REQUIRE(p && g && m && v && state, "fa_adam_step_state: bad argument");
  if (n == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_OPTIM, st);
  const bool al = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16) == 0;
  adam_kernel<<<ew_grid(n / 4 + 1), 256, 0, st>>>(p, g, m, v, n, beta1, beta2, eps, 0.f, 0.f, grad_scale, al ? 1 : 0, state);
  FA_LAUNCH_CHECK("fa_adam_step_state");
  return FA_OK;
}

int fa_sft_fuse_fwd(const float* x, const float* dcn, const float* gamma, const float* beta, float* out, int64_t n,
                    float slope, fa_stream_t stream) {
  FA_REQUIRE(x && dcn && gamma && beta && out, "fa_sft_fuse_fwd: null pointer");
  if (n == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  sft_fwd_kernel<<<ew_grid(n), 256, 0, st>>>(x, dcn, gamma, beta, out, n, slope);
  FA_LAUNCH_CHECK("fa_sft_fuse_fwd");
  return FA_OK;
}

int fa_sft_fuse_bwd(const float* x, const float* dcn, const float* gamma, const float* beta, const float* dout,
                    float* dx, float* ddcn, float* dgamma, float* dbeta, int64_t n, float slope, fa_stream_t stream) {
  FA_REQUIRE(x && dcn && gamma && beta && dout && dx && ddcn && dgamma && dbeta, "fa_sft_fuse_bwd: null pointer");
  if (n == 0) return FA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FaProfScope prof(FA_K_ELEMWISE, st);
  sft_bwd_kernel<<<ew_grid(n), 256, 0, st>>>(x, dcn, gamma, beta, dout, dx, ddcn, dgamma, dbeta, n, slope);
  FA_LAUNCH_CHECK("fa_sft_fuse_bwd");
  return FA_OK;
}

}  // extern "C"
