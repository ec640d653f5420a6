// Shared host/device helpers for the rsn_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "rsn_b200 kernels are written for sm_100a (B200) only"
#endif

// ---- error reporting (C-ABI: 0 ok, <0 bad argument, >0 cudaError_t) -------------------------
extern thread_local char g_rsn_err[512];
static inline int rsn_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_rsn_err, sizeof(g_rsn_err), fmt, ap);
  va_end(ap);
  return code;
}
#define RSN_ARG(cond, ...)                                  \
  do {                                                      \
    if (!(cond)) return rsn_fail(-1, __VA_ARGS__);          \
  } while (0)
#define RSN_LAUNCH_CHECK(name)                                                     \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess)                                                        \
      return rsn_fail((int)e__, "%s: %s", name, cudaGetErrorString(e__));          \
  } while (0)
#define RSN_CUDA(call)                                                             \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess)                                                        \
      return rsn_fail((int)e__, "%s: %s", #call, cudaGetErrorString(e__));         \
  } while (0)

static inline int rsn_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- device helpers ---------------------------------------------------------------------------
#define RSN_FULL 0xffffffffu

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(RSN_FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(RSN_FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
// torch.nan_to_num defaults: nan -> 0, +inf -> FLT_MAX, -inf -> -FLT_MAX
__device__ __forceinline__ float nan_to_num(float x, float nan_val = 0.f) {
  if (x != x) return nan_val;
  if (x == INFINITY) return 3.4028234663852886e38f;
  if (x == -INFINITY) return -3.4028234663852886e38f;
  return x;
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + __expf(-x)); }
// torch.nn.Softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplusf(float x) { return x > 20.f ? x : log1pf(expf(x)); }
