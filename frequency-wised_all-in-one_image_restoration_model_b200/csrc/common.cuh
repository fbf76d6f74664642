// Shared helpers for the freqair sm_100a kernels (device + host side of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define FA_OK 0
#define FA_ERR_ARG 1
#define FA_ERR_CUDA 2
#define FA_ERR_UNSUPPORTED 3

void fa_set_error(const char* fmt, ...);

#define FA_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) {                                              \
      fa_set_error(__VA_ARGS__);                                \
      return FA_ERR_ARG;                                        \
    }                                                           \
  } while (0)

#define FA_LAUNCH_CHECK(name)                                              \
  do {                                                                     \
    cudaError_t e__ = cudaGetLastError();                                  \
    if (e__ != cudaSuccess) {                                              \
      fa_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return FA_ERR_CUDA;                                                  \
    }                                                                      \
  } while (0)

#define FA_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      fa_set_error("%s failed: %s", #call, cudaGetErrorString(e__));          \
      return FA_ERR_CUDA;                                                     \
    }                                                                         \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize once per (call site, device): the attribute belongs to the device's copy of
// the function, so a process-wide flag would leave the kernel un-configured on every GPU but the first one used.
// Usage: FA_SMEM_ATTR_ONCE(bytes, kernel<template, args>);
#define FA_SMEM_ATTR_ONCE(bytes, ...)                                                                     \
  do {                                                                                                    \
    static bool done__[64] = {};                                                                          \
    int dev__ = 0;                                                                                        \
    FA_CUDA(cudaGetDevice(&dev__));                                                                       \
    if (dev__ < 0 || dev__ >= 64 || !done__[dev__]) {                                                     \
      FA_CUDA(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      if (dev__ >= 0 && dev__ < 64) done__[dev__] = true;                                                 \
    }                                                                                                     \
  } while (0)

// ---- optional in-stream timing of one kernel class (bench.py roofline leg) ----
// Each public launcher names its class; when profiling of that class is on, the
// launcher brackets the launch with events on the launching stream.
enum FaKernelClass {
  FA_K_NONE = 0,
  FA_K_GEMM = 1,
  FA_K_WIN_ATTN = 2,
  FA_K_JOINT_ATTN = 3,
  FA_K_BAND_FILTER = 4,
  FA_K_LAYERNORM = 5,
  FA_K_DWCONV = 6,
  FA_K_IM2COL = 7,
  FA_K_BN = 8,
  FA_K_OPTIM = 9,
  FA_K_DCN = 10,
  FA_K_ELEMWISE = 11,
  FA_K_COUNT = 12
};
struct FaProfScope {
  int cls; cudaStream_t st; cudaEvent_t e0; bool on;
  FaProfScope(int cls_, cudaStream_t st_);
  ~FaProfScope();
};
void fa_count_launch(int cls);

#ifdef __CUDACC__
constexpr int kNumSMs = 148;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// activations shared by GEMM epilogues / conv kernels
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_LRELU = 2, ACT_SIGMOID = 3,
       ACT_MUL = 4 };   // aux_act only: aux already holds the derivative (gelu'(u) stored by the forward), multiply by it

// Exact-erf GELU (nn.GELU default) and its derivative.  Phi(x) = 0.5 (1 + erf(x / sqrt 2)) is evaluated as
// 0.5 erfc(|z|) = 0.5 poly5(t) exp(-z^2), t = 1 / (1 + p |z|)  (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7 on erf)
// and reflected for x >= 0.  One rcp, one ex2 and eight FMAs instead of the two-branch erff() plus a second
// exponential: the GELU epilogues of the K <= 224 contractions were bound by erff issue slots (0.25 -> 0.47 ms for
// leff1 at T = 262144), not by HBM.  Measured against float64: |gelu error| <= 4.2e-7, |gelu' error| <= 3.2e-7 on
// [-9, 9] - tighter than torch's own fp32 gelu (1.2e-6, from x * ulp at large x).  exp(-z^2) = exp(-x^2 / 2) is shared
// with the Gaussian term of the derivative.
__device__ __forceinline__ float gelu_cdf(float x, float& e) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  e = __expf(-z * z);
  const float h = 0.5f * p * t * e;
  return x >= 0.f ? 1.0f - h : h;
}
__device__ __forceinline__ float gelu_f(float x) {
  float e;
  return x * gelu_cdf(x, e);
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float e;
  const float cdf = gelu_cdf(x, e);
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}
// gelu(x) and gelu'(x) from one cdf / exp evaluation
__device__ __forceinline__ float gelu_pair_f(float x, float& dg) {
  float e;
  const float cdf = gelu_cdf(x, e);
  dg = fmaf(x * 0.39894228040143267794f, e, cdf);
  return x * cdf;
}
__device__ __forceinline__ float act_f(float x, int act, float p) {
  if (act == ACT_GELU) return gelu_f(x);
  if (act == ACT_LRELU) return x > 0.f ? x : x * p;
  if (act == ACT_SIGMOID) return 1.0f / (1.0f + __expf(-x));
  return x;
}
__device__ __forceinline__ float act_grad_f(float x, int act, float p) {   // d act / d x at pre-activation x
  if (act == ACT_GELU) return gelu_grad_f(x);
  if (act == ACT_LRELU) return x > 0.f ? 1.0f : p;
  if (act == ACT_SIGMOID) { float s = 1.0f / (1.0f + __expf(-x)); return s * (1.0f - s); }
  if (act == ACT_MUL) return x;
  return 1.0f;
}
#endif
