// Shared host/device helpers for the rsn_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>

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

// SM count of the CURRENT device (cached per device id; the cache is immutable after its first, idempotent fill).
static inline int rsn_num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::atomic<int>& slot = cache[dev & 63];
  int n = slot.load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    slot.store(n, std::memory_order_relaxed);
  }
  return n;
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device function attribute: set it once per device (bit `dev` of
// `done`); two threads racing on first use both set the same value.
template <class K>
static inline cudaError_t rsn_ensure_smem(K kernel, int bytes, std::atomic<unsigned long long>& done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
  return e;
}
// Experiment switches (alternative kernel forms, timing ablations whose results are wrong by design) exist only in the
// test build (-DRSN_DEBUG_SWITCHES = librsn_b200_dbg.so); the product library never reads the environment.
static inline int rsn_env_int(const char* name, int dflt) {
#ifdef RSN_DEBUG_SWITCHES
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
#else
  (void)name;
  return dflt;
#endif
}
// Number of valid rays of a launch: the host capacity `cap`, or min(cap, *dev_count) when the count lives on the device
// (bounce passes: the number of masked rays never visits the host).
__device__ __forceinline__ int64_t rsn_count(int64_t cap, const int* dev_count) {
  if (dev_count == nullptr) return cap;
  const int64_t m = (int64_t)__ldg(dev_count);
  return m < cap ? (m < 0 ? 0 : m) : cap;
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
