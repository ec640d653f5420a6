// Stand-alone encoders behind the reference's component API (reflect_sampling_nerf_b200/components.py) and the direct
// parity tests of SURVEY.md §8 rows a7 / a11:
//   rsn_ipe_encode   NeRFEncoding(in_dim=3, num_frequencies=16, min_freq_exp=0, max_freq_exp=16, include_input=True)
//                    .forward(x, covs) (reflect_sampling_nerf_model.py:98-100): [P,3],[P,3,3] -> [P,99]
//   rsn_ide_encode   IntegratedSHEncoding.forward(directions, roughness) (reflect_sampling_nerf_components.py:52-140):
//                    [P,3],[P] -> [P,34]
// Same device functions as the fused field kernel's prologue / epilogue (csrc/encodings.cuh), fp32 out instead of the
// bf16 operand image.  Elementwise, HBM-bound (396 B / 136 B out per point); not on the model's path.
#include "encodings.cuh"

namespace {
using namespace rsnenc;

__global__ void __launch_bounds__(256) ipe_encode_kernel(const float* __restrict__ x, const float* __restrict__ cov,
                                                         float* __restrict__ out, int64_t n) {
  const int64_t pt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= n) return;
  float* o = out + pt * 99;
#pragma unroll 1
  for (int a = 0; a < 3; ++a) {
    const float xa = __ldg(x + pt * 3 + a);
    const float sx = __fmul_rn(6.2831854820251465f, xa);
    const float va = cov ? __ldg(cov + pt * 9 + a * 4) : 0.f;
    for (int k = 0; k < 16; ++k) {
      o[a * 16 + k] = ipe_value(sx, va, c_freq[k], 0);
      o[48 + a * 16 + k] = ipe_value(sx, va, c_freq[k], 1);
    }
    o[96 + a] = xa;
  }
}

__global__ void __launch_bounds__(256) ide_encode_kernel(const float* __restrict__ dirs, const float* __restrict__ rho,
                                                         float* __restrict__ out, int64_t n) {
  const int64_t pt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= n) return;
  const float d[3] = {__ldg(dirs + pt * 3), __ldg(dirs + pt * 3 + 1), __ldg(dirs + pt * 3 + 2)};
  float t[48];
  ide_features(d, __ldg(rho + pt), t);
#pragma unroll
  for (int i = 0; i < 34; ++i) out[pt * 34 + i] = t[i];
}
}  // namespace

extern "C" int rsn_ipe_encode(const float* x, const float* cov, float* out99, int64_t n_points, cudaStream_t stream) {
  RSN_ARG(n_points >= 0, "rsn_ipe_encode: bad shape");
  if (n_points == 0) return 0;
  RSN_ARG(x && out99, "rsn_ipe_encode: null pointer");
  ipe_encode_kernel<<<(unsigned)((n_points + 255) / 256), 256, 0, stream>>>(x, cov, out99, n_points);
  RSN_LAUNCH_CHECK("ipe_encode_kernel");
  return 0;
}

extern "C" int rsn_ide_encode(const float* dirs, const float* roughness, float* out34, int64_t n_points,
                              cudaStream_t stream) {
  RSN_ARG(n_points >= 0, "rsn_ide_encode: bad shape");
  if (n_points == 0) return 0;
  RSN_ARG(dirs && roughness && out34, "rsn_ide_encode: null pointer");
  ide_encode_kernel<<<(unsigned)((n_points + 255) / 256), 256, 0, stream>>>(dirs, roughness, out34, n_points);
  RSN_LAUNCH_CHECK("ide_encode_kernel");
  return 0;
}
