// K9: reflection set-up, device-side compaction of the bouncing rays, the reflected ray bundle and the composition of the
// bounce -- SURVEY.md §2.4 (K9), §8 rows a15, a16, a18.
//
// Replaces the per-ray eager ops of reflect_sampling_nerf_model.py:215-229 (renderer_rgb / renderer_factor /
// renderer_normals outputs of the fine pass, n.d, mask), 267-289 (boolean-mask gathers, origins, reflected directions,
// sqradius, the reflected RayBundle) and 240-241 / 311-313 / 337-339 (white (1 - acc) fallback, masked composition
// clip(diff + tint * (rgb + bg (1 - acc)))).  All per-ray and HBM-trivial; the point is
//   * a handful of launches instead of ~60 tiny ones, and
//   * NO HOST SYNC: the reference's `x[mask, :]` indexing reads the number of masked rays M back to the host
//     (model.py:229,267-289).  Here the mask is compacted on the device (rsn_reflect_compact: ascending ray order, the
//     order boolean indexing produces), M stays in a device int, and every kernel of the bounce passes takes that count
//     through its `n_rays_dev` argument while its buffers are sized for the capacity N.
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

// Stream compaction of the mask by ONE block: thread t owns rays [t c, (t+1) c), counts its masked rays, the block scans
// the 1024 counts, and each thread writes the ascending ray indices of its masked rays (idx) and every ray's rank among
// the masked rays or -1 (inv).  N is a ray batch (<= a few 100k): ~10 us, and it replaces a device->host read of M.
constexpr int COMPACT_THREADS = 1024;
__global__ void __launch_bounds__(COMPACT_THREADS) reflect_compact_kernel(const uint8_t* __restrict__ mask,
                                                                          int64_t* __restrict__ idx, int32_t* __restrict__ inv,
                                                                          int32_t* __restrict__ count, int64_t n) {
  __shared__ int warp_tot[COMPACT_THREADS / 32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t c = (n + COMPACT_THREADS - 1) / COMPACT_THREADS;
  const int64_t r0 = min((int64_t)t * c, n), r1 = min(r0 + c, n);
  int mine = 0;
  for (int64_t r = r0; r < r1; ++r) mine += __ldg(mask + r) ? 1 : 0;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(RSN_FULL, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(RSN_FULL, v, o);
      if (lane >= o) v += u;
    }
    warp_tot[lane] = v;   // inclusive totals of the warps
  }
  __syncthreads();
  int j = incl - mine + (warp > 0 ? warp_tot[warp - 1] : 0);
  for (int64_t r = r0; r < r1; ++r) {
    if (__ldg(mask + r)) {
      idx[j] = r;
      inv[r] = j++;
    } else {
      inv[r] = -1;
    }
  }
  if (t == COMPACT_THREADS - 1) *count = warp_tot[COMPACT_THREADS / 32 - 1];
}

// Reflected ray bundle of the masked rays (model.py:267-289): row j < M = ray idx[j].
//   sqradius = 2 |n.d| roughness^2 (model.py:272), pixel_area = pi sqradius (model.py:286); roughness = comp16[:, 12]
__global__ void __launch_bounds__(256) reflect_bundle_fwd_kernel(
    const int64_t* __restrict__ idx, const int32_t* __restrict__ count, const float* __restrict__ o2_all,
    const float* __restrict__ wr_all, const float* __restrict__ ndd, const float* __restrict__ comp16,
    float* __restrict__ o2, float* __restrict__ wr, float* __restrict__ sqr, float* __restrict__ area, int64_t n) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= rsn_count(n, count)) return;
  const int64_t r = __ldg(idx + j);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    o2[j * 3 + i] = __ldg(o2_all + r * 3 + i);
    wr[j * 3 + i] = __ldg(wr_all + r * 3 + i);
  }
  const float rough = __ldg(comp16 + r * 16 + 12);
  const float s = __fmul_rn(__fmul_rn(2.0f, fabsf(__ldg(ndd + r))), __fmul_rn(rough, rough));
  sqr[j] = s;
  area[j] = __fmul_rn(3.1415927410125732f, s);
}
// d/d roughness of the above, as a full gradient of the composited feature rows: g_comp16[r] = 0 except column 12.
__global__ void __launch_bounds__(256) reflect_bundle_bwd_kernel(
    const int32_t* __restrict__ inv, const float* __restrict__ ndd, const float* __restrict__ comp16,
    const float* __restrict__ g_sqr, const float* __restrict__ g_area, float* __restrict__ g_comp16, int64_t n) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int j = __ldg(inv + r);
  float g = 0.f;
  if (j >= 0) {
    const float gs = (g_sqr ? __ldg(g_sqr + j) : 0.f) + (g_area ? 3.1415927410125732f * __ldg(g_area + j) : 0.f);
    g = gs * 4.f * fabsf(__ldg(ndd + r)) * __ldg(comp16 + r * 16 + 12);
  }
  float4* o = reinterpret_cast<float4*>(g_comp16 + r * 16);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  o[0] = z, o[1] = z, o[2] = z, o[3] = make_float4(g, 0.f, 0.f, 0.f);
}

// out[r] = white (1 - acc_fine[r]) for rays that do not bounce (model.py:240-241), else
//          clip(diff[r] + tint[r] * refl, 0, 1), refl = comp_rgb[j] + bg[j] (1 - acc_r[j]), j = inv[r]   (model.py:311-313)
// (refl itself clamped to [0,1] in eval mode: RGBRenderer).  comp has row stride comp_ld (>= 3).  Optionally scatters the
// bounce's median depth to the ray grid (0 where no bounce): the padded form of outputs["depth_reflect_fine"].
__global__ void __launch_bounds__(256) reflect_compose_fwd_kernel(
    const float* __restrict__ acc_fine, const float* __restrict__ diff, const float* __restrict__ tint,
    const int32_t* __restrict__ inv, const float* __restrict__ comp, int64_t comp_ld, const float* __restrict__ bg,
    const float* __restrict__ acc_r, int clamp_inner, float* __restrict__ out, const float* __restrict__ depth_r,
    float* __restrict__ depth_out, int64_t n) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int j = __ldg(inv + r);
  if (j < 0) {
    const float v = 1.f - __ldg(acc_fine + r);
    out[r * 3 + 0] = v, out[r * 3 + 1] = v, out[r * 3 + 2] = v;
    if (depth_out) depth_out[r] = 0.f;
    return;
  }
  const float one_m = 1.f - __ldg(acc_r + j);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float refl = __ldg(comp + (int64_t)j * comp_ld + c) + __ldg(bg + (int64_t)j * 3 + c) * one_m;
    if (clamp_inner) refl = fminf(fmaxf(refl, 0.f), 1.f);
    const float v = __ldg(diff + r * 3 + c) + __ldg(tint + r * 3 + c) * refl;
    out[r * 3 + c] = fminf(fmaxf(v, 0.f), 1.f);
  }
  if (depth_out) depth_out[r] = __ldg(depth_r + j);
}

// Backward of the above (training: clamp_inner = 0).  Rays that bounce: g_comp16[j] = (g, 0 ...) with
// g = g_out tint inside the clip, g_bg[j] = g (1 - acc_r[j]), g_acc_fine[r] = 0 (the row was overwritten);
// others: g_acc_fine[r] = -sum_c g_out[r, c].
__global__ void __launch_bounds__(256) reflect_compose_bwd_kernel(
    const float* __restrict__ g_out, const float* __restrict__ diff, const float* __restrict__ tint,
    const int32_t* __restrict__ inv, const float* __restrict__ comp, int64_t comp_ld, const float* __restrict__ bg,
    const float* __restrict__ acc_r, float* __restrict__ g_comp16, float* __restrict__ g_bg,
    float* __restrict__ g_acc_fine, int64_t n) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int j = __ldg(inv + r);
  const float go[3] = {__ldg(g_out + r * 3), __ldg(g_out + r * 3 + 1), __ldg(g_out + r * 3 + 2)};
  if (j < 0) {
    g_acc_fine[r] = -(go[0] + go[1] + go[2]);
    return;
  }
  g_acc_fine[r] = 0.f;
  const float one_m = 1.f - __ldg(acc_r + j);
  float g[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float refl = __ldg(comp + (int64_t)j * comp_ld + c) + __ldg(bg + (int64_t)j * 3 + c) * one_m;
    const float tn = __ldg(tint + r * 3 + c);
    const float v = __ldg(diff + r * 3 + c) + tn * refl;
    g[c] = (v >= 0.f && v <= 1.f) ? go[c] * tn : 0.f;   // torch.clip: gradient inside [0, 1]
    g_bg[(int64_t)j * 3 + c] = g[c] * one_m;
  }
  float4* o = reinterpret_cast<float4*>(g_comp16 + (int64_t)j * 16);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  o[0] = make_float4(g[0], g[1], g[2], 0.f), o[1] = z, o[2] = z, o[3] = z;
}

inline unsigned blocks_for(int64_t n) { return (unsigned)((n + 255) / 256); }

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
  reflect_setup_kernel<<<blocks_for(n_rays), 256, 0, stream>>>(comp16, acc, depth, origins, dirs, clamp01, diff, tint, normal,
                                                               n_dot_d, mask, bounce_origins, bounce_dirs, n_rays);
  RSN_LAUNCH_CHECK("reflect_setup_kernel");
  return 0;
}

extern "C" int rsn_reflect_compact(const uint8_t* mask, int64_t* idx, int32_t* inv, int32_t* count, int64_t n_rays,
                                   cudaStream_t stream) {
  RSN_ARG(n_rays >= 0 && n_rays < (int64_t)1 << 31, "rsn_reflect_compact: bad shape");
  RSN_ARG(count != nullptr, "rsn_reflect_compact: null pointer");
  if (n_rays == 0) {
    RSN_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), stream));
    return 0;
  }
  RSN_ARG(mask && idx && inv, "rsn_reflect_compact: null pointer");
  reflect_compact_kernel<<<1, COMPACT_THREADS, 0, stream>>>(mask, idx, inv, count, n_rays);
  RSN_LAUNCH_CHECK("reflect_compact_kernel");
  return 0;
}

extern "C" int rsn_reflect_bundle_fwd(const int64_t* idx, const int32_t* count, const float* bounce_origins_all,
                                      const float* bounce_dirs_all, const float* n_dot_d, const float* comp16,
                                      float* origins, float* dirs, float* sqradius, float* pixel_area, int64_t n_rays,
                                      cudaStream_t stream) {
  RSN_ARG(n_rays >= 0, "rsn_reflect_bundle_fwd: bad shape");
  if (n_rays == 0) return 0;
  RSN_ARG(idx && count && bounce_origins_all && bounce_dirs_all && n_dot_d && comp16 && origins && dirs && sqradius &&
              pixel_area, "rsn_reflect_bundle_fwd: null pointer");
  reflect_bundle_fwd_kernel<<<blocks_for(n_rays), 256, 0, stream>>>(idx, count, bounce_origins_all, bounce_dirs_all, n_dot_d,
                                                                    comp16, origins, dirs, sqradius, pixel_area, n_rays);
  RSN_LAUNCH_CHECK("reflect_bundle_fwd_kernel");
  return 0;
}

extern "C" int rsn_reflect_bundle_bwd(const int32_t* inv, const float* n_dot_d, const float* comp16,
                                      const float* grad_sqradius, const float* grad_pixel_area, float* grad_comp16,
                                      int64_t n_rays, cudaStream_t stream) {
  RSN_ARG(n_rays >= 0, "rsn_reflect_bundle_bwd: bad shape");
  if (n_rays == 0) return 0;
  RSN_ARG(inv && n_dot_d && comp16 && grad_comp16, "rsn_reflect_bundle_bwd: null pointer");
  RSN_ARG(((uintptr_t)grad_comp16 & 15) == 0, "rsn_reflect_bundle_bwd: grad_comp16 must be 16-byte aligned");
  reflect_bundle_bwd_kernel<<<blocks_for(n_rays), 256, 0, stream>>>(inv, n_dot_d, comp16, grad_sqradius, grad_pixel_area,
                                                                    grad_comp16, n_rays);
  RSN_LAUNCH_CHECK("reflect_bundle_bwd_kernel");
  return 0;
}

extern "C" int rsn_reflect_compose_fwd(const float* acc_fine, const float* diff, const float* tint, const int32_t* inv,
                                       const float* comp, int64_t comp_ld, const float* bg, const float* acc_r,
                                       int clamp_inner, float* out, const float* depth_r, float* depth_out,
                                       int64_t n_rays, cudaStream_t stream) {
  RSN_ARG(n_rays >= 0 && comp_ld >= 3, "rsn_reflect_compose_fwd: bad shape");
  if (n_rays == 0) return 0;
  RSN_ARG(acc_fine && diff && tint && inv && comp && bg && acc_r && out, "rsn_reflect_compose_fwd: null pointer");
  RSN_ARG(!depth_out || depth_r, "rsn_reflect_compose_fwd: depth_r required with depth_out");
  reflect_compose_fwd_kernel<<<blocks_for(n_rays), 256, 0, stream>>>(acc_fine, diff, tint, inv, comp, comp_ld, bg, acc_r,
                                                                     clamp_inner, out, depth_r, depth_out, n_rays);
  RSN_LAUNCH_CHECK("reflect_compose_fwd_kernel");
  return 0;
}

extern "C" int rsn_reflect_compose_bwd(const float* grad_out, const float* diff, const float* tint, const int32_t* inv,
                                       const float* comp, int64_t comp_ld, const float* bg, const float* acc_r,
                                       float* grad_comp16, float* grad_bg, float* grad_acc_fine, int64_t n_rays,
                                       cudaStream_t stream) {
  RSN_ARG(n_rays >= 0 && comp_ld >= 3, "rsn_reflect_compose_bwd: bad shape");
  if (n_rays == 0) return 0;
  RSN_ARG(grad_out && diff && tint && inv && comp && bg && acc_r && grad_comp16 && grad_bg && grad_acc_fine,
          "rsn_reflect_compose_bwd: null pointer");
  RSN_ARG(((uintptr_t)grad_comp16 & 15) == 0, "rsn_reflect_compose_bwd: grad_comp16 must be 16-byte aligned");
  reflect_compose_bwd_kernel<<<blocks_for(n_rays), 256, 0, stream>>>(grad_out, diff, tint, inv, comp, comp_ld, bg, acc_r,
                                                                     grad_comp16, grad_bg, grad_acc_fine, n_rays);
  RSN_LAUNCH_CHECK("reflect_compose_bwd_kernel");
  return 0;
}
