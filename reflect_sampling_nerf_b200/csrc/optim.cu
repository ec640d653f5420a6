// Fused optimizer step of the field: RAdam (lr 1e-3, eps 1e-15) with the exponential learning-rate decay the reference
// configures (reflect_sampling_nerf_config.py:50-53: RAdamOptimizerConfig + ExponentialDecaySchedulerConfig(lr_final=1e-4,
// max_steps=50000)) over all 32 trained parameters in ONE launch -- SURVEY.md §8 row f2.
//
// Replaces torch.optim.RAdam's ~15 multi_tensor_apply launches + the scheduler's host arithmetic.  The step counter lives
// on the device (CUDA-graph capturable: nothing about the step number is baked into the launch), the data-parallel
// 1 / world_size averaging of the all-reduced gradient is folded in (`grad_scale`), and the arithmetic follows torch's
// foreach implementation (torch/optim/radam.py, _multi_tensor_radam) operation by operation in fp32 with the per-step
// scalars computed in fp64, so 100 steps stay within 1e-6 of torch.optim.RAdam (tests/test_optim_gpu.py).
// The bf16 operand re-pack that must follow is rsn_pack_field (csrc/pack.cu).
#include "rsn_common.cuh"
#include <math.h>
#include <algorithm>

extern "C" int64_t rsn_field_flat_layout(int64_t* host_offsets32);

namespace {

constexpr int N_PARAMS = 32;
constexpr int OPT_THREADS = 256;

struct RAdamParams {
  float* params[N_PARAMS];     // fp32 parameters in rsn_pack_field order
  int offs[N_PARAMS + 1];      // flat offsets (rsn_field_flat_layout)
  const float* grad;           // flat gradient vector
  float* exp_avg;              // flat
  float* exp_avg_sq;           // flat
  long long* step;             // device: optimizer steps taken so far; incremented by this launch
  unsigned int* counter;       // zero before the first launch; left zero
  float* lr_out;               // device: the learning rate this step used (for logging), or NULL
  double lr_init, lr_final;    // lr_final <= 0: constant lr
  long long max_steps;
  double beta1, beta2, eps;
  float grad_scale;
  int total;
};

__global__ void __launch_bounds__(OPT_THREADS) radam_kernel(const __grid_constant__ RAdamParams p) {
  __shared__ float s_c2, s_u, s_w1, s_b2, s_1mb2, s_eps;
  __shared__ bool last;
  if (threadIdx.x == 0) {
    const long long done = *p.step;          // every block reads the same value: the increment happens in the last block
    const double t = (double)(done + 1);
    double lr = p.lr_init;
    if (p.lr_final > 0.0 && p.max_steps > 0) {   // LambdaLR: the k-th optimizer step sees the schedule at k - 1
      double x = (double)done / (double)p.max_steps;
      x = x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x);
      lr = exp(log(p.lr_init) * (1.0 - x) + log(p.lr_final) * x);
    }
    const double bc1 = 1.0 - pow(p.beta1, t);
    const double b2t = pow(p.beta2, t);
    const double bc2 = 1.0 - b2t;
    const double rho_inf = 2.0 / (1.0 - p.beta2) - 1.0;
    const double rho_t = rho_inf - 2.0 * t * b2t / bc2;
    const double rect = rho_t > 5.0 ? sqrt((rho_t - 4.0) * (rho_t - 2.0) * rho_inf / ((rho_inf - 4.0) * (rho_inf - 2.0) * rho_t)) : 0.0;
    const double unrect = rect > 0.0 ? 0.0 : 1.0;
    s_u = (float)((lr * unrect / bc1) * -1.0);
    s_c2 = (float)(sqrt(bc2) * (lr * rect / bc1) * -1.0);
    s_w1 = (float)(1.0 - p.beta1);
    s_b2 = (float)p.beta2;
    s_1mb2 = (float)(1.0 - p.beta2);
    s_eps = (float)p.eps;
    if (blockIdx.x == 0 && p.lr_out) *p.lr_out = (float)lr;
  }
  __syncthreads();
  const float c2 = s_c2, u = s_u, w1 = s_w1, b2 = s_b2, omb2 = s_1mb2, eps = s_eps;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < p.total; e += gridDim.x * blockDim.x) {
    int lo = 0, hi = N_PARAMS;     // parameter k with offs[k] <= e < offs[k + 1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (p.offs[mid] <= e) lo = mid; else hi = mid;
    }
    float* const w = p.params[lo] + (e - p.offs[lo]);
    const float g = __ldg(p.grad + e) * p.grad_scale;
    float m = p.exp_avg[e], v = p.exp_avg_sq[e];
    m = fmaf(w1, g - m, m);                          // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(omb2 * g, g, v * b2);                   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    p.exp_avg[e] = m;
    p.exp_avg_sq[e] = v;
    float buf = sqrtf(v) + eps;                      // _foreach_sqrt, _foreach_add_(eps)
    buf = buf / c2;                                  // _foreach_div_(bias_correction2 * step size), 0 -> inf when unrectified
    buf = 1.0f / buf;                                // _foreach_reciprocal_
    buf = buf + u;                                   // _foreach_add_(unrect_step_size)
    *w = fmaf(m, buf, *w);                           // _foreach_addcmul_(params, exp_avgs, buffer)
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(p.counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *p.step += 1;
    *p.counter = 0u;
  }
}

}  // namespace

extern "C" int rsn_radam_step(float* const* params32, const float* flat_grad, float* exp_avg, float* exp_avg_sq,
                              long long* step_dev, unsigned int* counter_dev, float* lr_out_dev, double lr_init,
                              double lr_final, long long max_steps, double beta1, double beta2, double eps,
                              float grad_scale, cudaStream_t stream) {
  RSN_ARG(params32 && flat_grad && exp_avg && exp_avg_sq && step_dev && counter_dev, "rsn_radam_step: null pointer");
  RSN_ARG(lr_init > 0.0 && beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0,
          "rsn_radam_step: bad hyper-parameters");
  RAdamParams p = {};
  int64_t offs[N_PARAMS];
  const int64_t total = rsn_field_flat_layout(offs);
  for (int i = 0; i < N_PARAMS; ++i) {
    RSN_ARG(params32[i] != nullptr, "rsn_radam_step: params[%d] is null", i);
    p.params[i] = params32[i];
    p.offs[i] = (int)offs[i];
  }
  p.offs[N_PARAMS] = (int)total;
  p.grad = flat_grad;
  p.exp_avg = exp_avg;
  p.exp_avg_sq = exp_avg_sq;
  p.step = step_dev;
  p.counter = counter_dev;
  p.lr_out = lr_out_dev;
  p.lr_init = lr_init, p.lr_final = lr_final, p.max_steps = max_steps;
  p.beta1 = beta1, p.beta2 = beta2, p.eps = eps;
  p.grad_scale = grad_scale;
  p.total = (int)total;
  const int blocks = std::min((p.total + OPT_THREADS - 1) / OPT_THREADS, rsn_num_sms() * 8);
  radam_kernel<<<blocks, OPT_THREADS, 0, stream>>>(p);
  RSN_LAUNCH_CHECK("radam_kernel");
  return 0;
}
