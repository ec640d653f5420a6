// C-ABI housekeeping entry points (include/rsn_b200.h).
#include "rsn_common.cuh"

thread_local char g_rsn_err[512] = {0};

extern "C" const char* rsn_last_error(void) { return g_rsn_err; }
extern "C" int rsn_version(void) { return 200; }  // 0.2.0
extern "C" int rsn_device_ok(void) {
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess)
    return rsn_fail(1, "rsn_device_ok: no CUDA device");
  if (p.major != 10) return rsn_fail(-1, "rsn_device_ok: sm_%d%d is not sm_100 (B200)", p.major, p.minor);
  return 0;
}
