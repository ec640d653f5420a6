// K9: reflection set-up and composition of the bounce -- SURVEY.md §2.4 (K9), §8 rows a15-a18.
//
// Replaces the per-ray eager ops of reflect_sampling_nerf_model.py:215-229 (renderer_rgb / renderer_factor /
// renderer_normals outputs of the fine pass, n.d, mask), 267-272 (origins, reflected directions, sqradius) and
// 311-313 / 337-339 (masked composition clip(diff + tint * (rgb + bg (1 - acc)))).  All per-ray, HBM-trivial:
// the point is one launch instead of ~15 (set-up) and ~10 (composition, forward and backward) tiny ones.
// The boolean compaction itself (which rays bounce) stays a torch.nonzero on the mask this kernel writes.
#include "rsn_common.cuh"
#include "field_layout.cuh"

namespace {

// comp [N,16] = composited feature row of the fine pass (rsnf feature layout), acc [N], depth [N] (median).
__global__ void __launch_bounds__(256) reflect_setup_kernel(
    const float* __restrict__ comp, const float* __restrict__ acc, const float* __restrict__ depth,
    const float* __restrict__ origins, const float* __restrict__ dirs, int clamp01, float* __restrict__ diff,
    float* __restrict__ tint, float* __restrict__ nrm, float* __restrict__ ndd_out, uint8_t* __restrict__ mask,
    float* __restrict__ o2, float* __restrict__ wr, int64_t n) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float4* c4 = reinterpret_cast<const float4*>(comp + r * 16);
  const float4 c0 = __ldg(c4), c1 = __ldg(c4 + 1), c2 = __ldg(c4 + 2);
  const float a = __ldg(acc + r);
  const float o[3] = {__ldg(origins + r * 3), __ldg(origins + r * 3 + 1), __ldg(origins + r * 3 + 2)};
  const float d[3] = {__ldg(dirs + r * 3), __ldg(dirs + r * 3 + 1), __ldg(dirs + r * 3 + 2)};
  // renderer_rgb(diff, w): white background; renderer_factor(tint, w): unblended ("random"); eval clamps (App. A.6)
  float df[3] = {c0.w + (1.f - a), c1.x + (1.f - a), c1.y + (1.f - a)};
  float tn[3] = {c1.z, c1.w, c2.x};
  if (clamp01) {
#pragma unroll
    for (int i = 0; i < 3; ++i) df[i] = fminf(fmaxf(df[i], 0.f), 1.f), tn[i] = fminf(fmaxf(tn[i], 0.f), 1.f);
  }
  // NormalsRenderer: safe_normalize(sum w n) = v / (|v| + 1e-10)
  const float v[3] = {c2.y, c2.z, c2.w};
  const float inv = 1.f / (sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) + 1e-10f);
  const float nn[3] = {v[0] * inv, v[1] * inv, v[2] * inv};
  const float ndd = nn[0] * d[0] + nn[1] * d[1] + nn[2] * d[2];           // model.py:222
  // reflected direction: normalize(d - 2 (n.d) n)  (model.py:269-270; F.normalize eps 1e-12)
  float w[3] = {d[0] - 2.f * ndd * nn[0], d[1] - 2.f * ndd * nn[1], d[2] - 2.f * ndd * nn[2]};
  const float wl = fmaxf(sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]), 1e-12f);
  const float dep = __ldg(depth + r);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    diff[r * 3 + i] = df[i];
    tint[r * 3 + i] = tn[i];
    nrm[r * 3 + i] = nn[i];
    o2[r * 3 + i] = o[i] + dep * d[i];                                      // model.py:267
    wr[r * 3 + i] = w[i] / wl;
  }
  ndd_out[r] = ndd;
  mask[r] = (a > 1e-2f && ndd < 0.f) ? 1 : 0;                               // model.py:229
}

// out = base; out[idx[m]] = clip(diff[idx] + tint[idx] * refl, 0, 1), refl = comp_rgb[m] + bg[m] (1 - acc[m])
// (refl itself clamped to [0,1] in eval mode: RGBRenderer).  comp has row stride comp_ld (>= 3).
__global__ void __launch_bounds__(256) reflect_compose_fwd_kernel(
    const float* __restrict__ base, const float* __restrict__ diff, const float* __restrict__ tint,
    const int64_t* __restrict__ idx, const float* __restrict__ comp, int64_t comp_ld, const float* __restrict__ bg,
    const float* __restrict__ acc, int clamp_inner, float* __restrict__ out, int64_t n, int64_t m) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * 3) return;
  const int64_t j = t / 3;
  const int c = (int)(t - j * 3);
  const int64_t r = __ldg(idx + j);
  float refl = __ldg(comp + j * comp_ld + c) + __ldg(bg + j * 3 + c) * (1.f - __ldg(acc + j));
  if (clamp_inner) refl = fminf(fmaxf(refl, 0.f), 1.f);
  const float v = __ldg(diff + r * 3 + c) + __ldg(tint + r * 3 + c) * refl;
  out[r * 3 + c] = fminf(fmaxf(v, 0.f), 1.f);
}

// g_comp[m,3], g_bg[m,3] from g_out[N,3]; g_base = g_out with the bounced rows zeroed (they were overwritten).
__global__ void __launch_bounds__(256) reflect_compose_bwd_kernel(
    const float* __restrict__ g_out, const float* __restrict__ diff, const float* __restrict__ tint,
    const int64_t* __restrict__ idx, const float* __restrict__ comp, int64_t comp_ld, const float* __restrict__ bg,
    const float* __restrict__ acc, float* __restrict__ g_comp, float* __restrict__ g_bg, float* __restrict__ g_base,
    int64_t m) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * 3) return;
  const int64_t j = t / 3;
  const int c = (int)(t - j * 3);
  const int64_t r = __ldg(idx + j);
  const float one_m = 1.f - __ldg(acc + j);
  const float refl = __ldg(comp + j * comp_ld + c) + __ldg(bg + j * 3 + c) * one_m;
  const float tn = __ldg(tint + r * 3 + c);
  const float v = __ldg(diff + r * 3 + c) + tn * refl;
  const float g = (v >= 0.f && v <= 1.f) ? __ldg(g_out + r * 3 + c) * tn : 0.f;   // torch.clip: grad inside [0, 1]
  g_comp[j * 3 + c] = g;
  g_bg[j * 3 + c] = g * one_m;
  g_base[r * 3 + c] = 0.f;
}

}  // namespace

extern "C" int rsn_reflect_setup(const float* comp16, const float* acc, const float* depth, const float* origins,
                                 const float* dirs, int clamp01, float* diff, float* tint, float* normal, float* n_dot_d,
                                 uint8_t* mask, float* bounce_origins, float* bounce_dirs, int64_t n_rays,
                                 cudaStream_t stream) {
  RSN_ARG(n_rays >= 0, "rsn_reflect_setup: bad shape");
  if (n_rays == 0) return 0;
  RSN_ARG(comp16 && acc && depth && origins && dirs && diff && tint && normal && n_dot_d && mask && bounce_origins &&
              bounce_dirs, "rsn_reflect_setup: null pointer");
  RSN_ARG(((uintptr_t)comp16 & 15) == 0, "rsn_reflect_setup: comp16 must be 16-byte aligned");
  reflect_setup_kernel<<<(unsigned)((n_rays + 255) / 256), 256, 0, stream>>>(comp16, acc, depth, origins, dirs, clamp01, diff,
                                                                             tint, normal, n_dot_d, mask, bounce_origins,
                                                                             bounce_dirs, n_rays);
  RSN_LAUNCH_CHECK("reflect_setup_kernel");
  return 0;
}

extern "C" int rsn_reflect_compose_fwd(const float* base, const float* diff, const float* tint, const int64_t* idx,
                                       const float* comp, int64_t comp_ld, const float* bg, const float* acc,
                                       int clamp_inner, float* out, int64_t n_rays, int64_t n_bounced, cudaStream_t stream) {
  RSN_ARG(n_rays >= 0 && n_bounced >= 0 && comp_ld >= 3, "rsn_reflect_compose_fwd: bad shape");
  RSN_ARG(out && (n_rays == 0 || base), "rsn_reflect_compose_fwd: null pointer");
  if (n_rays == 0) return 0;
  if (out != base) RSN_CUDA(cudaMemcpyAsync(out, base, (size_t)n_rays * 3 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  if (n_bounced == 0) return 0;
  RSN_ARG(diff && tint && idx && comp && bg && acc, "rsn_reflect_compose_fwd: null pointer");
  reflect_compose_fwd_kernel<<<(unsigned)((n_bounced * 3 + 255) / 256), 256, 0, stream>>>(base, diff, tint, idx, comp, comp_ld, bg,
                                                                                        acc, clamp_inner, out, n_rays, n_bounced);
  RSN_LAUNCH_CHECK("reflect_compose_fwd_kernel");
  return 0;
}

extern "C" int rsn_reflect_compose_bwd(const float* grad_out, const float* diff, const float* tint, const int64_t* idx,
                                       const float* comp, int64_t comp_ld, const float* bg, const float* acc,
                                       float* grad_comp, float* grad_bg, float* grad_base, int64_t n_rays,
                                       int64_t n_bounced, cudaStream_t stream) {
  RSN_ARG(n_rays >= 0 && n_bounced >= 0 && comp_ld >= 3, "rsn_reflect_compose_bwd: bad shape");
  if (n_rays == 0) return 0;
  RSN_ARG(grad_out && grad_base, "rsn_reflect_compose_bwd: null pointer");
  if (grad_base != grad_out)
    RSN_CUDA(cudaMemcpyAsync(grad_base, grad_out, (size_t)n_rays * 3 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  if (n_bounced == 0) return 0;
  RSN_ARG(diff && tint && idx && comp && bg && acc && grad_comp && grad_bg, "rsn_reflect_compose_bwd: null pointer");
  reflect_compose_bwd_kernel<<<(unsigned)((n_bounced * 3 + 255) / 256), 256, 0, stream>>>(grad_out, diff, tint, idx, comp, comp_ld,
                                                                                        bg, acc, grad_comp, grad_bg, grad_base,
                                                                                        n_bounced);
  RSN_LAUNCH_CHECK("reflect_compose_bwd_kernel");
  return 0;
}
