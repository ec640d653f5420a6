// Stand-alone forms of the two geometry steps the fused field kernel runs in its prologue, for the reference's
// method-by-method Field API (reflect_sampling_nerf_b200/field.py) and for direct parity tests of SURVEY.md §8 rows a5/a6:
//   rsn_frustum_gaussians  field.get_blob   (reflect_sampling_nerf_field.py:90-96 -> Frustums.get_gaussian_blob ->
//                          conical_frustum_to_gaussian): mean [P,3], FULL cov [P,3,3]
//   rsn_contract           field.contract   (field.py:98-119): mip-NeRF-360 contraction of the mean and J cov J with the
//                          ReLU on the diagonal
// The hot path never materialises these (a C2 pass would write 100 MB of covariances that only feed their own diagonal,
// App. B Q9); the model goes through csrc/field_fwd.cu.  HBM-bound elementwise kernels, 48 B out per point.
#include "rsn_common.cuh"

namespace {

__global__ void __launch_bounds__(256) frustum_gaussians_kernel(const float* __restrict__ origins, const float* __restrict__ dirs,
                                                                const float* __restrict__ area, const float* __restrict__ bins,
                                                                int S, float* __restrict__ mean, float* __restrict__ cov,
                                                                int64_t n_points) {
  const int64_t pt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= n_points) return;
  const int64_t ray = pt / S;
  const int s = (int)(pt - ray * S);
  float o[3], d[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) o[a] = __ldg(origins + ray * 3 + a), d[a] = __ldg(dirs + ray * 3 + a);
  const float t0 = __ldg(bins + ray * (S + 1) + s), t1 = __ldg(bins + ray * (S + 1) + s + 1);
  // same operation order as frustum_gaussian_contracted in csrc/field_fwd.cu
  const float radius = __fdiv_rn(__fsqrt_rn(__ldg(area + ray)), 1.7724538509055159f);
  const float mu = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
  const float hw = __fmul_rn(__fsub_rn(t1, t0), 0.5f);
  const float hw2 = __fmul_rn(hw, hw), mu2 = __fmul_rn(mu, mu);
  const float den = __fadd_rn(__fmul_rn(3.0f, mu2), hw2);
  const float tmean = __fadd_rn(mu, __fdiv_rn(__fmul_rn(__fmul_rn(2.0f, mu), hw2), den));
#pragma unroll
  for (int a = 0; a < 3; ++a) mean[pt * 3 + a] = __fadd_rn(o[a], __fmul_rn(d[a], tmean));
  const float hw4 = hw2 * hw2;
  const float dir_var = hw2 / 3.0f - 0.26666668f * ((hw4 * (12.0f * mu2 - hw2)) / (den * den));
  const float rad_var = (radius * radius) * (mu2 * 0.25f + 0.41666666f * hw2 - 0.26666668f * hw4 / den);
  const float dd = fmaxf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2], 1e-10f);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      cov[pt * 9 + i * 3 + j] = dir_var * (d[i] * d[j]) + rad_var * ((i == j ? 1.0f : 0.0f) - d[i] * (d[j] / dd));
}

__global__ void __launch_bounds__(256) contract_kernel(const float* __restrict__ mean, const float* __restrict__ cov,
                                                       float* __restrict__ mean_out, float* __restrict__ cov_out,
                                                       int64_t n_points) {
  const int64_t pt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= n_points) return;
  float m[3], c[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    m[i] = __ldg(mean + pt * 3 + i);
#pragma unroll
    for (int j = 0; j < 3; ++j) c[i][j] = __ldg(cov + pt * 9 + i * 3 + j);
  }
  const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(m[0], m[0]), __fmul_rn(m[1], m[1])), __fmul_rn(m[2], m[2]));
  const float n1 = __fsqrt_rn(n2);
  float J[3][3];
  const bool outside = n1 > 1.0f;
  const float sc = outside ? __fdiv_rn(__fsub_rn(__fmul_rn(2.0f, n1), 1.0f), n2) : 1.0f;
  const float a2 = 2.0f * n1 - 2.0f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float e = (i == j) ? 1.0f : 0.0f;
      J[i][j] = outside ? (a2 * (e - m[i] * m[j] / n2) + e) / n2 : e;
    }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    mean_out[pt * 3 + i] = outside ? __fmul_rn(sc, m[i]) : m[i];
    float t[3];   // (J cov)_i.
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 3; ++j) s += J[i][j] * c[j][k];
      t[k] = s;
    }
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) s += t[k] * J[k][l];
      cov_out[pt * 9 + i * 3 + l] = (i == l) ? fmaxf(s, 0.f) : s;   // in-place ReLU on the diagonal (field.py:114-115)
    }
  }
}

}  // namespace

extern "C" int rsn_frustum_gaussians(const float* origins, const float* dirs, const float* pixel_area, const float* bins,
                                     int64_t n_rays, int64_t n_samples, float* mean, float* cov, cudaStream_t stream) {
  RSN_ARG(n_rays >= 0 && n_samples >= 1, "rsn_frustum_gaussians: bad shape");
  if (n_rays == 0) return 0;
  RSN_ARG(origins && dirs && pixel_area && bins && mean && cov, "rsn_frustum_gaussians: null pointer");
  const int64_t n = n_rays * n_samples;
  frustum_gaussians_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(origins, dirs, pixel_area, bins, (int)n_samples, mean,
                                                                            cov, n);
  RSN_LAUNCH_CHECK("frustum_gaussians_kernel");
  return 0;
}

extern "C" int rsn_contract(const float* mean, const float* cov, float* mean_out, float* cov_out, int64_t n_points,
                            cudaStream_t stream) {
  RSN_ARG(n_points >= 0, "rsn_contract: bad shape");
  if (n_points == 0) return 0;
  RSN_ARG(mean && cov && mean_out && cov_out, "rsn_contract: null pointer");
  contract_kernel<<<(unsigned)((n_points + 255) / 256), 256, 0, stream>>>(mean, cov, mean_out, cov_out, n_points);
  RSN_LAUNCH_CHECK("contract_kernel");
  return 0;
}
