// The eight loss terms of ReflectSamplingNeRFModel.get_loss_dict (reflect_sampling_nerf_model.py:395-429) in one launch
// (+ one for the backward) -- SURVEY.md §8 row a19.
//
//   k = 0..3  MSELoss(image, pred_k)  pred = mid_rgb_coarse, mid_rgb_fine, mid_reflect_coarse, mid_reflect_fine   (model.py:395-401)
//   k = 4..7  sum over rays of the per-ray sums the compositing kernel produced (csrc/composite.cu):
//             predicted_normal_loss_coarse / _fine, orientation_loss_coarse / _fine                                 (model.py:403-407)
// each scaled by its coefficient (misc.scale_dict, model.py:429; the pipeline's warm-up rewrites four of them every step,
// reflect_sampling_nerf_pipeline.py:79-91, so the coefficients are a DEVICE vector: the launch stays CUDA-graph capturable).
// out[0..7] = the scaled terms, out[8] = their sum (what the trainer back-propagates).
//
// blend_background_for_loss_computation is the identity for the white tensor background and an RGB target (SURVEY.md
// App. A.6), and the reference's .item() prints (model.py:409-410, host syncs) are not reproduced (App. B Q12).
// HBM-trivial (60 B per ray); deterministic: per-block partial sums in a workspace, summed in block order by the last
// block to finish.
#include "rsn_common.cuh"
#include <algorithm>

namespace {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_MAX_BLOCKS = 512;

struct LossParams {
  const float* pred[4];   // [N,3]
  const float* image;     // [N,3]
  const float* sums[4];   // [N] or NULL: pnl_coarse, pnl_fine, ol_coarse, ol_fine
  const float* coef;      // [8] device
  float* out;             // [9]
  double* partials;       // [LOSS_MAX_BLOCKS][8]
  unsigned int* counter;  // zero before the first launch; the last block leaves it zero again
  int64_t n;
};

__global__ void __launch_bounds__(LOSS_THREADS) loss_fwd_kernel(const __grid_constant__ LossParams p) {
  __shared__ double red[LOSS_THREADS / 32][8];
  __shared__ bool last;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < p.n; r += (int64_t)gridDim.x * blockDim.x) {
    const float im[3] = {__ldg(p.image + r * 3), __ldg(p.image + r * 3 + 1), __ldg(p.image + r * 3 + 2)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = im[c] - __ldg(p.pred[k] + r * 3 + c);
        s += d * d;
      }
      acc[k] += (double)s;
      if (p.sums[k]) acc[4 + k] += (double)__ldg(p.sums[k] + r);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(RSN_FULL, acc[k], o);
    if (lane == 0) red[warp][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double s = 0.0;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) s += red[w][threadIdx.x];
    p.partials[(size_t)blockIdx.x * 8 + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(p.counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < 8) {
    const int k = threadIdx.x;
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += p.partials[(size_t)b * 8 + k];
    if (k < 4) s /= (double)(3 * p.n);   // MSELoss: mean over N x 3
    red[0][k] = (double)((float)s * __ldg(p.coef + k));
    p.out[k] = (float)red[0][k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int k = 0; k < 8; ++k) tot += (float)red[0][k];   // the trainer's functools.reduce(torch.add, loss_dict.values())
    p.out[8] = tot;
    *p.counter = 0u;
  }
}

struct LossBwdParams {
  const float* pred[4];
  const float* image;
  const float* coef;
  const float* g_terms;   // [8] or NULL
  const float* g_total;   // [1] or NULL
  float* g_pred[4];       // [N,3]
  float* g_sums[4];       // [N] or NULL
  int64_t n;
};

__global__ void __launch_bounds__(LOSS_THREADS) loss_bwd_kernel(const __grid_constant__ LossBwdParams p) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.n) return;
  const float gt = p.g_total ? __ldg(p.g_total) : 0.f;
  const float inv = 2.f / (float)(3 * p.n);
  const float im[3] = {__ldg(p.image + r * 3), __ldg(p.image + r * 3 + 1), __ldg(p.image + r * 3 + 2)};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float g = ((p.g_terms ? __ldg(p.g_terms + k) : 0.f) + gt) * __ldg(p.coef + k) * inv;
#pragma unroll
    for (int c = 0; c < 3; ++c) p.g_pred[k][r * 3 + c] = g * (__ldg(p.pred[k] + r * 3 + c) - im[c]);
    if (p.g_sums[k]) p.g_sums[k][r] = ((p.g_terms ? __ldg(p.g_terms + 4 + k) : 0.f) + gt) * __ldg(p.coef + 4 + k);
  }
}

}  // namespace

extern "C" int64_t rsn_loss_workspace_bytes(void) { return (int64_t)LOSS_MAX_BLOCKS * 8 * sizeof(double) + 16; }

extern "C" int rsn_loss_fwd(const float* mid_rgb_coarse, const float* mid_rgb_fine, const float* mid_reflect_coarse,
                            const float* mid_reflect_fine, const float* image, const float* pnl_coarse,
                            const float* pnl_fine, const float* ol_coarse, const float* ol_fine, const float* coef8,
                            float* out9, void* workspace, int64_t n_rays, cudaStream_t stream) {
  RSN_ARG(n_rays >= 1, "rsn_loss_fwd: bad shape");
  RSN_ARG(mid_rgb_coarse && mid_rgb_fine && mid_reflect_coarse && mid_reflect_fine && image && coef8 && out9 && workspace,
          "rsn_loss_fwd: null pointer");
  RSN_ARG(((uintptr_t)workspace & 15) == 0, "rsn_loss_fwd: workspace must be 16-byte aligned");
  LossParams p = {};
  p.pred[0] = mid_rgb_coarse, p.pred[1] = mid_rgb_fine, p.pred[2] = mid_reflect_coarse, p.pred[3] = mid_reflect_fine;
  p.image = image;
  p.sums[0] = pnl_coarse, p.sums[1] = pnl_fine, p.sums[2] = ol_coarse, p.sums[3] = ol_fine;
  p.coef = coef8;
  p.out = out9;
  p.counter = (unsigned int*)workspace;
  p.partials = (double*)((uint8_t*)workspace + 16);
  p.n = n_rays;
  const int blocks = (int)std::min<int64_t>((n_rays + LOSS_THREADS - 1) / LOSS_THREADS, LOSS_MAX_BLOCKS);
  loss_fwd_kernel<<<blocks, LOSS_THREADS, 0, stream>>>(p);
  RSN_LAUNCH_CHECK("loss_fwd_kernel");
  return 0;
}

extern "C" int rsn_loss_bwd(const float* mid_rgb_coarse, const float* mid_rgb_fine, const float* mid_reflect_coarse,
                            const float* mid_reflect_fine, const float* image, const float* coef8, const float* grad_terms8,
                            const float* grad_total, float* g_mid_rgb_coarse, float* g_mid_rgb_fine,
                            float* g_mid_reflect_coarse, float* g_mid_reflect_fine, float* g_pnl_coarse, float* g_pnl_fine,
                            float* g_ol_coarse, float* g_ol_fine, int64_t n_rays, cudaStream_t stream) {
  RSN_ARG(n_rays >= 1, "rsn_loss_bwd: bad shape");
  RSN_ARG(mid_rgb_coarse && mid_rgb_fine && mid_reflect_coarse && mid_reflect_fine && image && coef8 && g_mid_rgb_coarse &&
              g_mid_rgb_fine && g_mid_reflect_coarse && g_mid_reflect_fine, "rsn_loss_bwd: null pointer");
  LossBwdParams p = {};
  p.pred[0] = mid_rgb_coarse, p.pred[1] = mid_rgb_fine, p.pred[2] = mid_reflect_coarse, p.pred[3] = mid_reflect_fine;
  p.image = image;
  p.coef = coef8;
  p.g_terms = grad_terms8;
  p.g_total = grad_total;
  p.g_pred[0] = g_mid_rgb_coarse, p.g_pred[1] = g_mid_rgb_fine, p.g_pred[2] = g_mid_reflect_coarse, p.g_pred[3] = g_mid_reflect_fine;
  p.g_sums[0] = g_pnl_coarse, p.g_sums[1] = g_pnl_fine, p.g_sums[2] = g_ol_coarse, p.g_sums[3] = g_ol_fine;
  p.n = n_rays;
  loss_bwd_kernel<<<(unsigned)((n_rays + LOSS_THREADS - 1) / LOSS_THREADS), LOSS_THREADS, 0, stream>>>(p);
  RSN_LAUNCH_CHECK("loss_bwd_kernel");
  return 0;
}
