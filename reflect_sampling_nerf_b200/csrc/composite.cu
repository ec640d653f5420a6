// K8: per-ray alpha compositing, forward and backward -- SURVEY.md §2.4, §8 rows a13/a14.
//
// Replaces (arithmetic restated in oracle/upstream.py):
//   RaySamples.get_weights           reflect_sampling_nerf_model.py:154,188,296,322
//   Accumulation/RGB/Depth(median)/Normals/Semantic renderers
//                                    reflect_sampling_nerf_model.py:155-156,176,189-190,210,215-226,311,337,341
//
// One warp per ray.  Lane l owns K consecutive samples [l*K, l*K+K) of a 32*K-sample round, so its
// loads are contiguous (float4 when the row allows it) and the transmittance scan is a K-step local
// prefix plus ONE 5-step warp scan of lane totals per round.  The running optical depth and the
// running weight sum (for the median) are carried in fp64: the reference's CPU cumsum accumulates in
// double, and exp(-tau) amplifies an fp32 scan error beyond the 1e-5 relative weight tolerance.
//
// HBM-bound.  Algorithmic bytes, forward: sigma 4 + start 4 + end 4 + C*4 feature in, weight 4 out per
// sample (28 B for C = 3) + (C + 2)*4 B out per ray.  Backward: those inputs again + 4 (dL/dw) in,
// 4 (dL/dsigma) + C*4 (dL/dfeat) out per sample.
#include "rsn_common.cuh"
#include <algorithm>
#include <initializer_list>

namespace {

template <int K>
struct Round {
  float dd[K], w[K];
};

// Shared forward math for one 32*K-sample round.  Returns weights in w[], updates carries.
template <int K>
__device__ __forceinline__ void weights_round(const float* __restrict__ sigma, const float* __restrict__ starts,
                                              const float* __restrict__ ends, int base, int S, int lane,
                                              double& tau_carry, float (&dd)[K], float (&w)[K],
                                              float (&t_next)[K]) {
  const int s0 = base + lane * K;
  double local = 0.0;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int s = s0 + k;
    float v = 0.f;
    if (s < S) v = (__ldg(ends + s) - __ldg(starts + s)) * __ldg(sigma + s);
    dd[k] = v;
    local += (double)v;
  }
  const double incl = warp_incl_scan(local, lane);
  double tau = tau_carry + (incl - local);  // optical depth in front of this lane's first sample
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float T = expf(-(float)tau);
    const float alpha = 1.f - expf(-dd[k]);
    w[k] = (s0 + k < S) ? nan_to_num(alpha * T) : 0.f;
    tau += (double)dd[k];
    t_next[k] = expf(-(float)tau);  // transmittance behind sample k (= d w_k / d dd_k)
  }
  tau_carry += __shfl_sync(RSN_FULL, incl, 31);
}

template <int K, int C>
__global__ void __launch_bounds__(256) composite_fwd_kernel(
    const float* __restrict__ sigma, const float* __restrict__ starts, const float* __restrict__ ends,
    int64_t bin_stride, const float* __restrict__ feat, float* __restrict__ weights, float* __restrict__ acc_out,
    float* __restrict__ depth_out, float* __restrict__ feat_out, int64_t n_rays, int S,
    const int* __restrict__ n_rays_dev) {
  n_rays = rsn_count(n_rays, n_rays_dev);
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n_rays; r += nwarps) {
    const float* sg = sigma + r * S;
    const float* st = starts + r * bin_stride;
    const float* en = ends + r * bin_stride;
    double tau_carry = 0.0, cw_carry = 0.0;
    float fsum[C > 0 ? C : 1];
#pragma unroll
    for (int c = 0; c < (C > 0 ? C : 1); ++c) fsum[c] = 0.f;
    float acc = 0.f;
    int median = S;  // searchsorted(cumsum(w), 0.5, side="left")
    for (int base = 0; base < S; base += 32 * K) {
      float dd[K], w[K], tn[K];
      weights_round<K>(sg, st, en, base, S, lane, tau_carry, dd, w, tn);
      const int s0 = base + lane * K;
      double local = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (s0 + k < S) {
          weights[r * S + s0 + k] = w[k];
          acc += w[k];
          if (C > 0) {
            const float* f = feat + ((int64_t)r * S + s0 + k) * C;
#pragma unroll
            for (int c = 0; c < C; ++c) fsum[c] += w[k] * __ldg(f + c);
          }
        }
        local += (double)w[k];
      }
      // median: first sample whose inclusive cumulative weight reaches 0.5
      const double incl = warp_incl_scan(local, lane);
      double cw = cw_carry + (incl - local);
      int mine = S;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        cw += (double)w[k];
        if (mine == S && s0 + k < S && (float)cw >= 0.5f) mine = s0 + k;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(RSN_FULL, mine, o));
      median = min(median, mine);
      cw_carry += __shfl_sync(RSN_FULL, incl, 31);
    }
    acc = warp_sum(acc);
#pragma unroll
    for (int c = 0; c < (C > 0 ? C : 1); ++c) fsum[c] = warp_sum(fsum[c]);
    if (lane == 0) {
      acc_out[r] = acc;
      const int mi = min(median, S - 1);
      depth_out[r] = (__ldg(st + mi) + __ldg(en + mi)) / 2.f;
      if (C > 0) {
#pragma unroll
        for (int c = 0; c < C; ++c) feat_out[r * C + c] = fsum[c];
      }
    }
  }
}

// Backward.  gw_s = dL/dw_s (explicit) + dL/dacc + sum_c dL/dfeat_out_c * feat_{s,c}
//            dL/d(dd_j) = gw_j * T_{j+1} - sum_{s>j} gw_s * w_s ;  dL/dsigma_j = dL/d(dd_j) * delta_j
// The suffix sum is total - inclusive prefix, so a first pass accumulates total = sum_s gw_s w_s.
template <int K, int C>
__global__ void __launch_bounds__(256) composite_bwd_kernel(
    const float* __restrict__ sigma, const float* __restrict__ starts, const float* __restrict__ ends,
    int64_t bin_stride, const float* __restrict__ feat, const float* __restrict__ g_weights,
    const float* __restrict__ g_acc, const float* __restrict__ g_feat_out, float* __restrict__ g_sigma,
    float* __restrict__ g_feat, int64_t n_rays, int S, const int* __restrict__ n_rays_dev) {
  n_rays = rsn_count(n_rays, n_rays_dev);
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n_rays; r += nwarps) {
    const float* sg = sigma + r * S;
    const float* st = starts + r * bin_stride;
    const float* en = ends + r * bin_stride;
    const float ga = g_acc ? __ldg(g_acc + r) : 0.f;
    float go[C > 0 ? C : 1];
#pragma unroll
    for (int c = 0; c < (C > 0 ? C : 1); ++c) go[c] = (C > 0 && g_feat_out) ? __ldg(g_feat_out + r * C + c) : 0.f;

    // pass 1: total = sum_s gw_s * w_s
    double tau_carry = 0.0;
    double total = 0.0;
    for (int base = 0; base < S; base += 32 * K) {
      float dd[K], w[K], tn[K];
      weights_round<K>(sg, st, en, base, S, lane, tau_carry, dd, w, tn);
      const int s0 = base + lane * K;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (s0 + k < S) {
          float gw = ga + (g_weights ? __ldg(g_weights + r * S + s0 + k) : 0.f);
          if (C > 0) {
            const float* f = feat + ((int64_t)r * S + s0 + k) * C;
#pragma unroll
            for (int c = 0; c < C; ++c) gw += go[c] * __ldg(f + c);
          }
          total += (double)(gw * w[k]);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(RSN_FULL, total, o);

    // pass 2: gradients
    tau_carry = 0.0;
    double pre_carry = 0.0;
    for (int base = 0; base < S; base += 32 * K) {
      float dd[K], w[K], tn[K], gw[K];
      weights_round<K>(sg, st, en, base, S, lane, tau_carry, dd, w, tn);
      const int s0 = base + lane * K;
      double local = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        gw[k] = 0.f;
        if (s0 + k < S) {
          float g = ga + (g_weights ? __ldg(g_weights + r * S + s0 + k) : 0.f);
          if (C > 0) {
            const float* f = feat + ((int64_t)r * S + s0 + k) * C;
            float* gf = g_feat ? g_feat + ((int64_t)r * S + s0 + k) * C : nullptr;
#pragma unroll
            for (int c = 0; c < C; ++c) {
              g += go[c] * __ldg(f + c);
              if (gf) gf[c] = go[c] * w[k];
            }
          }
          gw[k] = g;
        }
        local += (double)(gw[k] * w[k]);
      }
      const double incl = warp_incl_scan(local, lane);
      double pre = pre_carry + (incl - local);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        pre += (double)(gw[k] * w[k]);  // inclusive prefix through sample k
        if (s0 + k < S) {
          const float delta = __ldg(en + s0 + k) - __ldg(st + s0 + k);
          const float gdd = gw[k] * tn[k] - (float)(total - pre);
          g_sigma[r * S + s0 + k] = gdd * delta;
        }
      }
      pre_carry += __shfl_sync(RSN_FULL, incl, 31);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Vectorised forms for C = 4 Q channels (Q = 1, 2, 4; the model composites all 16 feature channels at once).
// The per-sample weights of a ray go through a warp-private shared-memory row; the features are then read as
// float4 with lane <-> (sample, channel quad), so one warp instruction covers 512 contiguous bytes, instead of
// one lane walking its samples' channels with 4-byte loads.  The backward makes a single pass over the features.
template <int K, int Q>
__global__ void __launch_bounds__(256) composite_fwd_vec_kernel(
    const float* __restrict__ sigma, const float* __restrict__ starts, const float* __restrict__ ends,
    int64_t bin_stride, const float4* __restrict__ feat4, float* __restrict__ weights, float* __restrict__ acc_out,
    float* __restrict__ depth_out, float4* __restrict__ feat_out4, int64_t n_rays, int S, int s_pad,
    const float* __restrict__ normals, float* __restrict__ pnl_out, float* __restrict__ ol_out,
    float* __restrict__ blend_out, const int* __restrict__ n_rays_dev) {
  extern __shared__ float sm_rows[];
  n_rays = rsn_count(n_rays, n_rays_dev);
  const int lane = threadIdx.x & 31;
  float* ws = sm_rows + (threadIdx.x >> 5) * s_pad;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n_rays; r += nwarps) {
    const float* sg = sigma + r * S;
    const float* st = starts + r * bin_stride;
    const float* en = ends + r * bin_stride;
    double tau_carry = 0.0, cw_carry = 0.0;
    float acc = 0.f;
    int median = S;
    for (int base = 0; base < S; base += 32 * K) {
      float dd[K], w[K], tn[K];
      weights_round<K>(sg, st, en, base, S, lane, tau_carry, dd, w, tn);
      const int s0 = base + lane * K;
      double local = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (s0 + k < S) {
          weights[r * S + s0 + k] = w[k];
          ws[s0 + k] = w[k];
          acc += w[k];
        }
        local += (double)w[k];
      }
      const double incl = warp_incl_scan(local, lane);
      double cw = cw_carry + (incl - local);
      int mine = S;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        cw += (double)w[k];
        if (mine == S && s0 + k < S && (float)cw >= 0.5f) mine = s0 + k;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(RSN_FULL, mine, o));
      median = min(median, mine);
      cw_carry += __shfl_sync(RSN_FULL, incl, 31);
    }
    acc = warp_sum(acc);
    __syncwarp();
    const float4* f4 = feat4 + (int64_t)r * S * Q;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    float lsum = 0.f;   // Q == 4 with normals: quad 2 accumulates w |n - n_pred|^2, quad 3 accumulates w max(0, n.d)^2
#pragma unroll 4
    for (int idx = lane; idx < S * Q; idx += 32) {
      const float4 v = __ldg(f4 + idx);
      const float wv = ws[idx / Q];
      a.x += wv * v.x, a.y += wv * v.y, a.z += wv * v.z, a.w += wv * v.w;
      if (Q == 4 && normals) {
        const int quad = lane & 3;
        if (quad == 2) {          // feature columns 8..11 = tint_b, pred_normal xyz
          const float* nn = normals + ((int64_t)r * S + idx / Q) * 3;
          const float d0 = __ldg(nn) - v.y, d1 = __ldg(nn + 1) - v.z, d2 = __ldg(nn + 2) - v.w;
          lsum += wv * (d0 * d0 + d1 * d1 + d2 * d2);
        } else if (quad == 3) {   // feature columns 12..15 = sigmoid roughness, n.d, raw density, softplus roughness
          const float pz = fmaxf(v.y, 0.f);
          lsum += wv * pz * pz;
        }
      }
    }
    if (Q == 4 && normals) {
#pragma unroll
      for (int o = 16; o >= 4; o >>= 1) lsum += __shfl_xor_sync(RSN_FULL, lsum, o);
      if (lane == 2) pnl_out[r] = lsum;
      if (lane == 3) ol_out[r] = lsum;
    }
#pragma unroll
    for (int o = 16; o >= Q; o >>= 1) {
      a.x += __shfl_xor_sync(RSN_FULL, a.x, o), a.y += __shfl_xor_sync(RSN_FULL, a.y, o);
      a.z += __shfl_xor_sync(RSN_FULL, a.z, o), a.w += __shfl_xor_sync(RSN_FULL, a.w, o);
    }
    if (lane < Q) feat_out4[r * Q + lane] = a;
    if (lane == 0) {
      if (blend_out) {   // RGBRenderer with the white background, then torch.clip (model.py:176-177): clip(rgb + (1 - acc))
        const float rest = 1.f - acc;
        blend_out[r * 3 + 0] = fminf(fmaxf(a.x + rest, 0.f), 1.f);
        blend_out[r * 3 + 1] = fminf(fmaxf(a.y + rest, 0.f), 1.f);
        blend_out[r * 3 + 2] = fminf(fmaxf(a.z + rest, 0.f), 1.f);
      }
      acc_out[r] = acc;
      const int mi = min(median, S - 1);
      depth_out[r] = (__ldg(st + mi) + __ldg(en + mi)) / 2.f;
    }
    __syncwarp();
  }
}

template <int K, int Q>
__global__ void __launch_bounds__(256) composite_bwd_vec_kernel(
    const float* __restrict__ sigma, const float* __restrict__ starts, const float* __restrict__ ends,
    int64_t bin_stride, const float4* __restrict__ feat4, const float* __restrict__ g_weights,
    const float* __restrict__ g_acc, const float4* __restrict__ g_feat_out4, float* __restrict__ g_sigma,
    float4* __restrict__ g_feat4, int64_t n_rays, int S, int s_pad, const float* __restrict__ normals,
    const float* __restrict__ g_pnl, const float* __restrict__ g_ol, const float* __restrict__ g_blend,
    const float* __restrict__ feat_out, const float* __restrict__ acc_in, const int* __restrict__ n_rays_dev) {
  extern __shared__ float sm_rows[];
  n_rays = rsn_count(n_rays, n_rays_dev);
  const int lane = threadIdx.x & 31;
  float* ws = sm_rows + (threadIdx.x >> 5) * 3 * s_pad;   // weights
  float* tns = ws + s_pad;                                // transmittance behind the sample
  float* gws = tns + s_pad;                               // dL/dw_s (explicit + accumulation + features)
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n_rays; r += nwarps) {
    const float* sg = sigma + r * S;
    const float* st = starts + r * bin_stride;
    const float* en = ends + r * bin_stride;
    float ga = g_acc ? __ldg(g_acc + r) : 0.f;
    float4 go = g_feat_out4 ? __ldg(g_feat_out4 + r * Q + (lane % Q)) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (g_blend) {   // blend = clip(feat_out[0:3] + (1 - acc), 0, 1): torch.clip passes the gradient inside [0, 1]
      const float rest = 1.f - __ldg(acc_in + r);
      float gb[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = __ldg(feat_out + r * (4 * Q) + c) + rest;
        gb[c] = (v >= 0.f && v <= 1.f) ? __ldg(g_blend + r * 3 + c) : 0.f;
      }
      ga -= gb[0] + gb[1] + gb[2];
      if ((lane % Q) == 0) go.x += gb[0], go.y += gb[1], go.z += gb[2];
    }
    double tau_carry = 0.0;
    for (int base = 0; base < S; base += 32 * K) {
      float dd[K], w[K], tn[K];
      weights_round<K>(sg, st, en, base, S, lane, tau_carry, dd, w, tn);
      const int s0 = base + lane * K;
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (s0 + k < S) ws[s0 + k] = w[k], tns[s0 + k] = tn[k];
    }
    __syncwarp();
    // one pass over the features: dL/dw_s and dL/dfeat
    const float4* f4 = feat4 + (int64_t)r * S * Q;
    float4* gf4 = g_feat4 ? g_feat4 + (int64_t)r * S * Q : nullptr;
#pragma unroll 4
    for (int idx0 = 0; idx0 < S * Q; idx0 += 32) {
      const int idx = idx0 + lane;
      const bool ok = idx < S * Q;
      const int smp = ok ? idx / Q : 0;
      const float4 v = ok ? __ldg(f4 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      float dot = go.x * v.x + go.y * v.y + go.z * v.z + go.w * v.w;
#pragma unroll
      for (int o = 1; o < Q; o <<= 1) dot += __shfl_xor_sync(RSN_FULL, dot, o);
      if (ok) {
        const float wv = ws[smp];
        if ((lane % Q) == 0) gws[smp] = ga + (g_weights ? __ldg(g_weights + r * S + smp) : 0.f) + dot;
        if (gf4) {
          float4 g = make_float4(go.x * wv, go.y * wv, go.z * wv, go.w * wv);
          if (Q == 4 && normals) {   // per-sample normal losses: the weight is a constant there (detached upstream)
            const int quad = lane & 3;
            if (quad == 2 && g_pnl) {
              const float* nn = normals + ((int64_t)r * S + smp) * 3;
              const float c = 2.f * wv * __ldg(g_pnl + r);
              g.y += c * (v.y - __ldg(nn)), g.z += c * (v.z - __ldg(nn + 1)), g.w += c * (v.w - __ldg(nn + 2));
            } else if (quad == 3 && g_ol) {
              g.y += 2.f * wv * __ldg(g_ol + r) * fmaxf(v.y, 0.f);
            }
          }
          gf4[idx] = g;
        }
      }
    }
    __syncwarp();
    if (g_sigma == nullptr) {   // density detached upstream (bounce passes, model.py:297,323)
      __syncwarp();
      continue;
    }
    // dL/d(dd_j) = gw_j T_{j+1} - sum_{s>j} gw_s w_s ;  dL/dsigma_j = dL/d(dd_j) * delta_j
    double total = 0.0;
    for (int s = lane; s < S; s += 32) total += (double)(gws[s] * ws[s]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(RSN_FULL, total, o);
    double pre_carry = 0.0;
    for (int base = 0; base < S; base += 32 * K) {
      const int s0 = base + lane * K;
      float gw[K], w[K];
      double local = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const bool ok = s0 + k < S;
        gw[k] = ok ? gws[s0 + k] : 0.f;
        w[k] = ok ? ws[s0 + k] : 0.f;
        local += (double)(gw[k] * w[k]);
      }
      const double incl = warp_incl_scan(local, lane);
      double pre = pre_carry + (incl - local);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        pre += (double)(gw[k] * w[k]);
        if (s0 + k < S) {
          const float delta = __ldg(en + s0 + k) - __ldg(st + s0 + k);
          g_sigma[r * S + s0 + k] = (gw[k] * tns[s0 + k] - (float)(total - pre)) * delta;
        }
      }
      pre_carry += __shfl_sync(RSN_FULL, incl, 31);
    }
    __syncwarp();
  }
}

// The upstream renderers' call forms (AccumulationRenderer / RGBRenderer / DepthRenderer(median) / NormalsRenderer /
// SemanticRenderer, reflect_sampling_nerf_model.py:117-124) take WEIGHTS, not densities: one warp per
// ray reduces acc = sum w, feat_out = sum w feat (C <= 16 channels) and the median depth from the given weights.
__global__ void __launch_bounds__(256) render_weights_kernel(const float* __restrict__ weights, const float* __restrict__ feat,
                                                             int C, const float* __restrict__ starts,
                                                             const float* __restrict__ ends, int64_t bin_stride,
                                                             float* __restrict__ acc_out, float* __restrict__ feat_out,
                                                             float* __restrict__ depth_out, int64_t n_rays, int S) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rays) return;
  const float* w = weights + r * S;
  float acc = 0.f, fs[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) fs[c] = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float ws = __ldg(w + s);
    acc += ws;
    for (int c = 0; c < C; ++c) fs[c] += ws * __ldg(feat + ((int64_t)r * S + s) * C + c);
  }
  acc = warp_sum(acc);
  if (lane == 0 && acc_out) acc_out[r] = acc;
  for (int c = 0; c < C; ++c) {
    const float v = warp_sum(fs[c]);
    if (lane == 0) feat_out[r * C + c] = v;
  }
  if (depth_out) {   // steps[clamp(searchsorted(cumsum(w), 0.5, side="left"), 0, S-1)], cumulative weight in fp64
    int median = S;
    double carry = 0.0;
    for (int base = 0; base < S; base += 32) {
      const int s = base + lane;
      const double v = s < S ? (double)__ldg(w + s) : 0.0;
      const double incl = warp_incl_scan(v, lane);
      int mine = (s < S && (float)(carry + incl) >= 0.5f) ? s : S;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(RSN_FULL, mine, o));
      median = min(median, mine);
      carry += __shfl_sync(RSN_FULL, incl, 31);
    }
    if (lane == 0) {
      const int mi = min(median, S - 1);
      depth_out[r] = (__ldg(starts + r * bin_stride + mi) + __ldg(ends + r * bin_stride + mi)) / 2.f;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The model's 16-channel form with the feature rows STAGED BY THE TMA ENGINE (S % 4 == 0, S <= 256).  A ray's features
// (S x 64 B) and normals (S x 12 B) are contiguous, so lane 0 of each warp fetches them with two cp.async.bulk copies
// into the warp's shared-memory stage (mbarrier complete_tx) and the warp computes the ray's weights from sigma / bins
// (small, register-staged loads) WHILE that copy is in flight; the feature pass then runs out of shared memory.  The
// register-staged kernels above wait on two dependent load phases per ray at 50 % occupancy (ncu: 23 warps stalled on
// the long scoreboard per issue, 33 % of the DRAM peak); here the two phases overlap and ~10 KB per warp are in flight
// without holding a register.  The feature pass is branch-free (the per-sample normal losses are selected per lane, not
// branched on).  Arithmetic and its order are those of the kernels above.
constexpr int TMA_WARPS = 4;
constexpr int TMA_MAX_SAMPLES = 256;

__device__ __forceinline__ uint32_t cs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cs_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cs_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void cs_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cs_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cs_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(cs_smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void cs_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   cs_smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(cs_smem_u32(bar))
               : "memory");
}
// lane 0: fetch ray r's features (and normals) into the warp's stage
__device__ __forceinline__ void cs_fetch_ray(float* stage, uint64_t* bar, const float* feat, const float* normals, int64_t r,
                                             int S) {
  cs_mbar_expect_tx(bar, (uint32_t)S * (normals ? 76u : 64u));
  cs_bulk_g2s(stage, feat + r * S * 16, (uint32_t)S * 64u, bar);
  if (normals) cs_bulk_g2s(stage + 16 * S, normals + r * S * 3, (uint32_t)S * 12u, bar);
}
// per-warp shared memory (floats): [feat 16 S | normals 3 S | work rows: forward S (weights), backward 3 S]
__host__ __device__ constexpr int cs_warp_floats(int S, bool bwd) { return 19 * S + (bwd ? 3 * S : S); }

template <int K>
__global__ void __launch_bounds__(TMA_WARPS * 32) composite16_fwd_tma_kernel(
    const float* __restrict__ sigma, const float* __restrict__ starts, const float* __restrict__ ends, int64_t bin_stride,
    const float* __restrict__ feat, float* __restrict__ weights, float* __restrict__ acc_out, float* __restrict__ depth_out,
    float4* __restrict__ feat_out4, int64_t n_rays, int S, const float* __restrict__ normals, float* __restrict__ pnl_out,
    float* __restrict__ ol_out, float* __restrict__ blend_out, const int* __restrict__ n_rays_dev) {
  extern __shared__ __align__(128) float sm_tma[];
  __shared__ uint64_t bars[TMA_WARPS];
  n_rays = rsn_count(n_rays, n_rays_dev);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* const fs = sm_tma + (size_t)warp * cs_warp_floats(S, false);
  float* const nrm = fs + 16 * S;
  float* const ws = nrm + 3 * S;
  uint64_t* const bar = &bars[warp];
  if (lane == 0) {
    cs_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int64_t warp0 = (int64_t)blockIdx.x * TMA_WARPS + warp, nwarps = (int64_t)gridDim.x * TMA_WARPS;
  uint32_t phase = 0;
  const int quad = lane & 3;
  const float sel2 = quad == 2 ? 1.f : 0.f, sel3 = quad == 3 ? 1.f : 0.f;
  if (warp0 < n_rays && lane == 0) cs_fetch_ray(fs, bar, feat, normals, warp0, S);
  for (int64_t r = warp0; r < n_rays; r += nwarps) {
    const float* sg = sigma + r * S;
    const float* st = starts + r * bin_stride;
    const float* en = ends + r * bin_stride;
    double tau_carry = 0.0, cw_carry = 0.0;
    float acc = 0.f;
    int median = S;
    for (int b0 = 0; b0 < S; b0 += 32 * K) {   // the ray's weights, while the TMA engine fetches its features
      float dd[K], w[K], tn[K];
      weights_round<K>(sg, st, en, b0, S, lane, tau_carry, dd, w, tn);
      const int s0 = b0 + lane * K;
      double local = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (s0 + k < S) {
          ws[s0 + k] = w[k];
          acc += w[k];
        }
        local += (double)w[k];
      }
      if (K == 4 && s0 + 3 < S) {
        *reinterpret_cast<float4*>(weights + r * S + s0) = make_float4(w[0], w[1], w[2], w[3]);
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k)
          if (s0 + k < S) weights[r * S + s0 + k] = w[k];
      }
      const double incl = warp_incl_scan(local, lane);
      double cw = cw_carry + (incl - local);
      int mine = S;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        cw += (double)w[k];
        if (mine == S && s0 + k < S && (float)cw >= 0.5f) mine = s0 + k;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(RSN_FULL, mine, o));
      median = min(median, mine);
      cw_carry += __shfl_sync(RSN_FULL, incl, 31);
    }
    acc = warp_sum(acc);
    __syncwarp();
    cs_mbar_wait(bar, phase);
    phase ^= 1u;
    const float4* const f4 = reinterpret_cast<const float4*>(fs);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    float lsum = 0.f;   // with normals: quad 2 accumulates w |n - n_pred|^2, quad 3 accumulates w max(0, n.d)^2
    if (normals) {
#pragma unroll 4
      for (int idx = lane; idx < S * 4; idx += 32) {
        const float4 v = f4[idx];
        const int smp = idx >> 2;
        const float wv = ws[smp];
        a.x += wv * v.x, a.y += wv * v.y, a.z += wv * v.z, a.w += wv * v.w;
        // quad 2: feature columns 8..11 = tint_b, pred_normal xyz; quad 3: columns 12..15 = sigmoid roughness, n.d, ...
        const float d0 = nrm[smp * 3] - v.y, d1 = nrm[smp * 3 + 1] - v.z, d2 = nrm[smp * 3 + 2] - v.w;
        const float pz = fmaxf(v.y, 0.f);
        lsum += wv * (sel2 * (d0 * d0 + d1 * d1 + d2 * d2) + sel3 * (pz * pz));
      }
#pragma unroll
      for (int o = 16; o >= 4; o >>= 1) lsum += __shfl_xor_sync(RSN_FULL, lsum, o);
      if (lane == 2) pnl_out[r] = lsum;
      if (lane == 3) ol_out[r] = lsum;
    } else {
#pragma unroll 4
      for (int idx = lane; idx < S * 4; idx += 32) {
        const float4 v = f4[idx];
        const float wv = ws[idx >> 2];
        a.x += wv * v.x, a.y += wv * v.y, a.z += wv * v.z, a.w += wv * v.w;
      }
    }
    __syncwarp();   // every lane is done with the stage and the work row: fetch the next ray
    if (r + nwarps < n_rays && lane == 0) cs_fetch_ray(fs, bar, feat, normals, r + nwarps, S);
#pragma unroll
    for (int o = 16; o >= 4; o >>= 1) {
      a.x += __shfl_xor_sync(RSN_FULL, a.x, o), a.y += __shfl_xor_sync(RSN_FULL, a.y, o);
      a.z += __shfl_xor_sync(RSN_FULL, a.z, o), a.w += __shfl_xor_sync(RSN_FULL, a.w, o);
    }
    if (lane < 4) feat_out4[r * 4 + lane] = a;
    if (lane == 0) {
      if (blend_out) {   // RGBRenderer with the white background, then torch.clip (model.py:176-177): clip(rgb + (1 - acc))
        const float rest = 1.f - acc;
        blend_out[r * 3 + 0] = fminf(fmaxf(a.x + rest, 0.f), 1.f);
        blend_out[r * 3 + 1] = fminf(fmaxf(a.y + rest, 0.f), 1.f);
        blend_out[r * 3 + 2] = fminf(fmaxf(a.z + rest, 0.f), 1.f);
      }
      acc_out[r] = acc;
      const int mi = min(median, S - 1);
      depth_out[r] = (__ldg(st + mi) + __ldg(en + mi)) / 2.f;
    }
  }
}

template <int K>
__global__ void __launch_bounds__(TMA_WARPS * 32) composite16_bwd_tma_kernel(
    const float* __restrict__ sigma, const float* __restrict__ starts, const float* __restrict__ ends, int64_t bin_stride,
    const float* __restrict__ feat, const float* __restrict__ g_weights, const float* __restrict__ g_acc,
    const float4* __restrict__ g_feat_out4, float* __restrict__ g_sigma, float4* __restrict__ g_feat4, int64_t n_rays, int S,
    const float* __restrict__ normals, const float* __restrict__ g_pnl, const float* __restrict__ g_ol,
    const float* __restrict__ g_blend, const float* __restrict__ feat_out, const float* __restrict__ acc_in,
    const int* __restrict__ n_rays_dev) {
  extern __shared__ __align__(128) float sm_tma[];
  __shared__ uint64_t bars[TMA_WARPS];
  n_rays = rsn_count(n_rays, n_rays_dev);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* const fs = sm_tma + (size_t)warp * cs_warp_floats(S, true);
  float* const nrm = fs + 16 * S;
  float* const ws = nrm + 3 * S;             // weights
  float* const tns = ws + S;                 // transmittance behind the sample
  float* const gws = tns + S;                // dL/dw_s (explicit + accumulation + features)
  uint64_t* const bar = &bars[warp];
  if (lane == 0) {
    cs_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int64_t warp0 = (int64_t)blockIdx.x * TMA_WARPS + warp, nwarps = (int64_t)gridDim.x * TMA_WARPS;
  uint32_t phase = 0;
  const int quad = lane & 3;
  if (warp0 < n_rays && lane == 0) cs_fetch_ray(fs, bar, feat, normals, warp0, S);
  for (int64_t r = warp0; r < n_rays; r += nwarps) {
    const float* sg = sigma + r * S;
    const float* st = starts + r * bin_stride;
    const float* en = ends + r * bin_stride;
    float ga = g_acc ? __ldg(g_acc + r) : 0.f;
    float4 go = g_feat_out4 ? __ldg(g_feat_out4 + r * 4 + quad) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (g_blend) {   // blend = clip(feat_out[0:3] + (1 - acc), 0, 1): torch.clip passes the gradient inside [0, 1]
      const float rest = 1.f - __ldg(acc_in + r);
      float gb[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = __ldg(feat_out + r * 16 + c) + rest;
        gb[c] = (v >= 0.f && v <= 1.f) ? __ldg(g_blend + r * 3 + c) : 0.f;
      }
      ga -= gb[0] + gb[1] + gb[2];
      if (quad == 0) go.x += gb[0], go.y += gb[1], go.z += gb[2];
    }
    // per-lane multipliers of the two normal-loss gradients (quad 2: 2 g_pnl; quad 3: 2 g_ol), no branches in the pass
    const float c2 = (normals && g_pnl && quad == 2) ? 2.f * __ldg(g_pnl + r) : 0.f;
    const float c3 = (normals && g_ol && quad == 3) ? 2.f * __ldg(g_ol + r) : 0.f;
    double tau_carry = 0.0;
    for (int b0 = 0; b0 < S; b0 += 32 * K) {   // weights, T_next and the explicit part of dL/dw while the features arrive
      float dd[K], w[K], tn[K];
      weights_round<K>(sg, st, en, b0, S, lane, tau_carry, dd, w, tn);
      const int s0 = b0 + lane * K;
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (s0 + k < S) {
          ws[s0 + k] = w[k], tns[s0 + k] = tn[k];
          gws[s0 + k] = ga + (g_weights ? __ldg(g_weights + r * S + s0 + k) : 0.f);
        }
    }
    __syncwarp();
    cs_mbar_wait(bar, phase);
    phase ^= 1u;
    // one pass over the features: dL/dw_s and dL/dfeat
    const float4* const f4 = reinterpret_cast<const float4*>(fs);
    float4* const gf4 = g_feat4 + (int64_t)r * S * 4;
#pragma unroll 4
    for (int idx = lane; idx < S * 4; idx += 32) {
      const int smp = idx >> 2;
      const float4 v = f4[idx];
      float dot = go.x * v.x + go.y * v.y + go.z * v.z + go.w * v.w;
      dot += __shfl_xor_sync(RSN_FULL, dot, 1);
      dot += __shfl_xor_sync(RSN_FULL, dot, 2);
      const float wv = ws[smp];
      if (quad == 0) gws[smp] += dot;
      float4 g = make_float4(go.x * wv, go.y * wv, go.z * wv, go.w * wv);
      if (normals) {   // per-sample normal losses: the weight is a constant there (detached upstream)
        const float k2 = c2 * wv, k3 = c3 * wv;
        g.y += k2 * (v.y - nrm[smp * 3]) + k3 * fmaxf(v.y, 0.f);
        g.z += k2 * (v.z - nrm[smp * 3 + 1]);
        g.w += k2 * (v.w - nrm[smp * 3 + 2]);
      }
      gf4[idx] = g;
    }
    __syncwarp();
    if (r + nwarps < n_rays && lane == 0) cs_fetch_ray(fs, bar, feat, normals, r + nwarps, S);
    if (g_sigma != nullptr) {   // (NULL: density detached upstream -- bounce passes, model.py:297,323)
      // dL/d(dd_j) = gw_j T_{j+1} - sum_{s>j} gw_s w_s ;  dL/dsigma_j = dL/d(dd_j) * delta_j
      double total = 0.0;
      for (int s = lane; s < S; s += 32) total += (double)(gws[s] * ws[s]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(RSN_FULL, total, o);
      double pre_carry = 0.0;
      for (int b0 = 0; b0 < S; b0 += 32 * K) {
        const int s0 = b0 + lane * K;
        float gw[K], w[K], out[K];
        double local = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const bool ok = s0 + k < S;
          gw[k] = ok ? gws[s0 + k] : 0.f;
          w[k] = ok ? ws[s0 + k] : 0.f;
          local += (double)(gw[k] * w[k]);
        }
        const double incl = warp_incl_scan(local, lane);
        double pre = pre_carry + (incl - local);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          pre += (double)(gw[k] * w[k]);
          out[k] = 0.f;
          if (s0 + k < S)
            out[k] = (gw[k] * tns[s0 + k] - (float)(total - pre)) * (__ldg(en + s0 + k) - __ldg(st + s0 + k));
        }
        if (K == 4 && s0 + 3 < S) {
          *reinterpret_cast<float4*>(g_sigma + r * S + s0) = make_float4(out[0], out[1], out[2], out[3]);
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k)
            if (s0 + k < S) g_sigma[r * S + s0 + k] = out[k];
        }
        pre_carry += __shfl_sync(RSN_FULL, incl, 31);
      }
    }
    __syncwarp();   // the work rows are free for the next ray
  }
}

// the TMA-staged form applies to the model's shapes: 16 channels, S a multiple of 4 up to 256, 16-byte aligned rows
inline bool tma_form_ok(int S, std::initializer_list<const void*> rows) {
  if (S % 4 != 0 || S > TMA_MAX_SAMPLES || S < 4) return false;
  for (const void* p : rows)
    if (((uintptr_t)p & 15) != 0) return false;
  return true;
}

constexpr int VEC_MAX_SAMPLES = 1024;   // 8 warps x 3 rows x 4 KB of dynamic shared memory in the backward

template <int C>
int launch_fwd(const float* sigma, const float* starts, const float* ends, int64_t bin_stride, const float* feat,
               float* weights, float* acc, float* depth, float* feat_out, int64_t n_rays, int S,
               cudaStream_t stream, const int* n_rays_dev = nullptr, const float* normals = nullptr, float* pnl = nullptr,
               float* ol = nullptr, float* blend = nullptr) {
  const int threads = 256;
  int64_t want = (n_rays * 32 + threads - 1) / threads;
  int blocks = (int)std::min<int64_t>(want, (int64_t)rsn_num_sms() * 16);
  if (C == 16 && tma_form_ok(S, {feat, feat_out, sigma, weights, normals})) {
    const size_t smem = (size_t)TMA_WARPS * cs_warp_floats(S, false) * sizeof(float);
    static std::atomic<unsigned long long> done[3];
    const int bpsm = std::max(1, std::min(8, (int)((225 * 1024) / (smem + 1024))));
    blocks = (int)std::min<int64_t>((n_rays + TMA_WARPS - 1) / TMA_WARPS, (int64_t)rsn_num_sms() * bpsm);
#define RSN_FWDT(K, I)                                                                                               \
  do {                                                                                                               \
    RSN_CUDA(rsn_ensure_smem(composite16_fwd_tma_kernel<K>, (int)(TMA_WARPS * cs_warp_floats(TMA_MAX_SAMPLES, false) * 4), done[I])); \
    composite16_fwd_tma_kernel<K><<<blocks, TMA_WARPS * 32, smem, stream>>>(sigma, starts, ends, bin_stride, feat, weights, \
                                                                          acc, depth, (float4*)feat_out, n_rays, S,    \
                                                                          normals, pnl, ol, blend, n_rays_dev);         \
  } while (0)
    if (S <= 32) RSN_FWDT(1, 0);
    else if (S <= 64) RSN_FWDT(2, 1);
    else RSN_FWDT(4, 2);
#undef RSN_FWDT
    RSN_LAUNCH_CHECK("composite16_fwd_tma_kernel");
    return 0;
  }
  if (C >= 4 && C % 4 == 0 && S <= VEC_MAX_SAMPLES && ((uintptr_t)feat & 15) == 0 && ((uintptr_t)feat_out & 15) == 0) {
    constexpr int Q = C >= 4 ? C / 4 : 1;
    const int s_pad = (S + 3) & ~3;
    const size_t smem = (size_t)(threads / 32) * s_pad * sizeof(float);
    blocks = (int)std::min<int64_t>(want, (int64_t)rsn_num_sms() * 8);
#define RSN_FWDV(K)                                                                                              \
  composite_fwd_vec_kernel<K, Q><<<blocks, threads, smem, stream>>>(sigma, starts, ends, bin_stride,             \
                                                                    (const float4*)feat, weights, acc, depth,    \
                                                                    (float4*)feat_out, n_rays, S, s_pad, normals, \
                                                                    pnl, ol, blend, n_rays_dev)
    if (S <= 32) RSN_FWDV(1);
    else if (S <= 64) RSN_FWDV(2);
    else RSN_FWDV(4);
#undef RSN_FWDV
    RSN_LAUNCH_CHECK("composite_fwd_vec_kernel");
    return 0;
  }
  if (normals || blend) return rsn_fail(-1, "rsn_composite16_fwd: needs n_samples <= %d and 16-byte aligned feat", VEC_MAX_SAMPLES);
#define RSN_FWD(K)                                                                                        \
  composite_fwd_kernel<K, C><<<blocks, threads, 0, stream>>>(sigma, starts, ends, bin_stride, feat, weights, \
                                                             acc, depth, feat_out, n_rays, S, n_rays_dev)
  if (S <= 32) RSN_FWD(1);
  else if (S <= 64) RSN_FWD(2);
  else RSN_FWD(4);
#undef RSN_FWD
  RSN_LAUNCH_CHECK("composite_fwd_kernel");
  return 0;
}

template <int C>
int launch_bwd(const float* sigma, const float* starts, const float* ends, int64_t bin_stride, const float* feat,
               const float* g_weights, const float* g_acc, const float* g_feat_out, float* g_sigma, float* g_feat,
               int64_t n_rays, int S, cudaStream_t stream, const int* n_rays_dev = nullptr, const float* normals = nullptr,
               const float* g_pnl = nullptr, const float* g_ol = nullptr, const float* g_blend = nullptr,
               const float* feat_out = nullptr, const float* acc_in = nullptr) {
  const int threads = 256;
  int64_t want = (n_rays * 32 + threads - 1) / threads;
  int blocks = (int)std::min<int64_t>(want, (int64_t)rsn_num_sms() * 16);
  if (C == 16 && g_feat && tma_form_ok(S, {feat, g_feat_out, g_feat, sigma, g_weights, g_sigma, normals})) {
    const size_t smem = (size_t)TMA_WARPS * cs_warp_floats(S, true) * sizeof(float);
    static std::atomic<unsigned long long> done_t[3];
    const int bpsm = std::max(1, std::min(8, (int)((225 * 1024) / (smem + 1024))));
    blocks = (int)std::min<int64_t>((n_rays + TMA_WARPS - 1) / TMA_WARPS, (int64_t)rsn_num_sms() * bpsm);
#define RSN_BWDT(K, I)                                                                                               \
  do {                                                                                                               \
    RSN_CUDA(rsn_ensure_smem(composite16_bwd_tma_kernel<K>, (int)(TMA_WARPS * cs_warp_floats(TMA_MAX_SAMPLES, true) * 4), done_t[I])); \
    composite16_bwd_tma_kernel<K><<<blocks, TMA_WARPS * 32, smem, stream>>>(                                          \
        sigma, starts, ends, bin_stride, feat, g_weights, g_acc, (const float4*)g_feat_out, g_sigma, (float4*)g_feat, \
        n_rays, S, normals, g_pnl, g_ol, g_blend, feat_out, acc_in, n_rays_dev);                                     \
  } while (0)
    if (S <= 32) RSN_BWDT(1, 0);
    else if (S <= 64) RSN_BWDT(2, 1);
    else RSN_BWDT(4, 2);
#undef RSN_BWDT
    RSN_LAUNCH_CHECK("composite16_bwd_tma_kernel");
    return 0;
  }
  if (C >= 4 && C % 4 == 0 && S <= VEC_MAX_SAMPLES && ((uintptr_t)feat & 15) == 0 && ((uintptr_t)g_feat_out & 15) == 0 &&
      ((uintptr_t)g_feat & 15) == 0) {
    constexpr int Q = C >= 4 ? C / 4 : 1;
    const int s_pad = (S + 3) & ~3;
    const size_t smem = (size_t)(threads / 32) * 3 * s_pad * sizeof(float);
    static std::atomic<unsigned long long> done[3];
    RSN_CUDA(rsn_ensure_smem(composite_bwd_vec_kernel<1, Q>, 8 * 3 * VEC_MAX_SAMPLES * 4, done[0]));
    RSN_CUDA(rsn_ensure_smem(composite_bwd_vec_kernel<2, Q>, 8 * 3 * VEC_MAX_SAMPLES * 4, done[1]));
    RSN_CUDA(rsn_ensure_smem(composite_bwd_vec_kernel<4, Q>, 8 * 3 * VEC_MAX_SAMPLES * 4, done[2]));
    blocks = (int)std::min<int64_t>(want, (int64_t)rsn_num_sms() * 8);
#define RSN_BWDV(K)                                                                                                \
  composite_bwd_vec_kernel<K, Q><<<blocks, threads, smem, stream>>>(sigma, starts, ends, bin_stride,               \
                                                                    (const float4*)feat, g_weights, g_acc,         \
                                                                    (const float4*)g_feat_out, g_sigma,            \
                                                                    (float4*)g_feat, n_rays, S, s_pad, normals,    \
                                                                    g_pnl, g_ol, g_blend, feat_out, acc_in, n_rays_dev)
    if (S <= 32) RSN_BWDV(1);
    else if (S <= 64) RSN_BWDV(2);
    else RSN_BWDV(4);
#undef RSN_BWDV
    RSN_LAUNCH_CHECK("composite_bwd_vec_kernel");
    return 0;
  }
  if (normals || g_blend || !g_sigma)
    return rsn_fail(-1, "rsn_composite16_bwd: needs n_samples <= %d and 16-byte aligned buffers", VEC_MAX_SAMPLES);
#define RSN_BWD(K)                                                                                          \
  composite_bwd_kernel<K, C><<<blocks, threads, 0, stream>>>(sigma, starts, ends, bin_stride, feat, g_weights, \
                                                             g_acc, g_feat_out, g_sigma, g_feat, n_rays, S, n_rays_dev)
  if (S <= 32) RSN_BWD(1);
  else if (S <= 64) RSN_BWD(2);
  else RSN_BWD(4);
#undef RSN_BWD
  RSN_LAUNCH_CHECK("composite_bwd_kernel");
  return 0;
}

}  // namespace

#define RSN_DISPATCH_C(FN, ...)                \
  switch (n_channels) {                        \
    case 0: return FN<0>(__VA_ARGS__);         \
    case 1: return FN<1>(__VA_ARGS__);         \
    case 3: return FN<3>(__VA_ARGS__);         \
    case 4: return FN<4>(__VA_ARGS__);         \
    case 8: return FN<8>(__VA_ARGS__);         \
    case 16: return FN<16>(__VA_ARGS__);       \
    default: return rsn_fail(-1, "rsn_composite: n_channels must be one of 0,1,3,4,8,16 (got %d)", (int)n_channels); \
  }

extern "C" int rsn_composite_fwd(const float* sigma, const float* starts, const float* ends, int64_t bin_row_stride,
                                 const float* feat, int64_t n_channels, float* weights, float* accumulation,
                                 float* depth_median, float* feat_out, int64_t n_rays, int64_t n_samples,
                                 const int* n_rays_dev, cudaStream_t stream) {
  RSN_ARG(n_rays >= 0 && n_samples >= 1, "rsn_composite_fwd: bad shape");
  if (n_rays == 0) return 0;
  RSN_ARG(sigma && starts && ends && weights && accumulation && depth_median, "rsn_composite_fwd: null pointer");
  RSN_ARG(n_channels == 0 || (feat && feat_out), "rsn_composite_fwd: feat/feat_out required when n_channels > 0");
  RSN_DISPATCH_C(launch_fwd, sigma, starts, ends, bin_row_stride, feat, weights, accumulation, depth_median,
                 feat_out, n_rays, (int)n_samples, stream, n_rays_dev);
}

extern "C" int rsn_composite_bwd(const float* sigma, const float* starts, const float* ends, int64_t bin_row_stride,
                                 const float* feat, int64_t n_channels, const float* grad_weights,
                                 const float* grad_accumulation, const float* grad_feat_out, float* grad_sigma,
                                 float* grad_feat, int64_t n_rays, int64_t n_samples, const int* n_rays_dev,
                                 cudaStream_t stream) {
  RSN_ARG(n_rays >= 0 && n_samples >= 1, "rsn_composite_bwd: bad shape");
  if (n_rays == 0) return 0;
  RSN_ARG(sigma && starts && ends && grad_sigma, "rsn_composite_bwd: null pointer");
  RSN_ARG(n_channels == 0 || feat, "rsn_composite_bwd: feat required when n_channels > 0");
  RSN_DISPATCH_C(launch_bwd, sigma, starts, ends, bin_row_stride, feat, grad_weights, grad_accumulation,
                 grad_feat_out, grad_sigma, grad_feat, n_rays, (int)n_samples, stream, n_rays_dev);
}

// The model's 16-channel form.  Optional riders on the same pass:
//   * normals != NULL: the two per-sample normal losses (their sample weights are the detached compositing weights,
//     reflect_sampling_nerf_model.py:403-407), per ray
//       pred_normal_loss[r] = sum_s w_s |normals_s - feat_s[9:12]|^2    orientation_loss[r] = sum_s w_s max(0, feat_s[13])^2
//   * rgb_blend != NULL: renderer_rgb's white-background blend and the clip that follows it (model.py:176-177,210-211),
//       rgb_blend[r] = clip(feat_out[r, 0:3] + (1 - accumulation[r]), 0, 1)
//   * n_rays_dev != NULL: the ray count lives on the device (bounce passes)
extern "C" int rsn_composite16_fwd(const float* sigma, const float* starts, const float* ends, int64_t bin_row_stride,
                                   const float* feat, const float* normals, float* weights, float* accumulation,
                                   float* depth_median, float* feat_out, float* pred_normal_loss,
                                   float* orientation_loss, float* rgb_blend, int64_t n_rays, int64_t n_samples,
                                   const int* n_rays_dev, cudaStream_t stream) {
  RSN_ARG(n_rays >= 0 && n_samples >= 1, "rsn_composite16_fwd: bad shape");
  if (n_rays == 0) return 0;
  RSN_ARG(sigma && starts && ends && feat && weights && accumulation && depth_median && feat_out,
          "rsn_composite16_fwd: null pointer");
  RSN_ARG(!normals || (pred_normal_loss && orientation_loss), "rsn_composite16_fwd: loss outputs required with normals");
  return launch_fwd<16>(sigma, starts, ends, bin_row_stride, feat, weights, accumulation, depth_median, feat_out, n_rays,
                        (int)n_samples, stream, n_rays_dev, normals, pred_normal_loss, orientation_loss, rgb_blend);
}

// grad_sigma == NULL: the density is detached upstream (bounce passes), only grad_feat is produced.
// grad_rgb_blend != NULL needs the forward's feat_out and accumulation (the clip mask is recomputed from them).
extern "C" int rsn_composite16_bwd(const float* sigma, const float* starts, const float* ends, int64_t bin_row_stride,
                                   const float* feat, const float* normals, const float* grad_weights,
                                   const float* grad_accumulation, const float* grad_feat_out,
                                   const float* grad_pred_normal_loss, const float* grad_orientation_loss,
                                   const float* grad_rgb_blend, const float* feat_out, const float* accumulation,
                                   float* grad_sigma, float* grad_feat, int64_t n_rays, int64_t n_samples,
                                   const int* n_rays_dev, cudaStream_t stream) {
  RSN_ARG(n_rays >= 0 && n_samples >= 1, "rsn_composite16_bwd: bad shape");
  if (n_rays == 0) return 0;
  RSN_ARG(sigma && starts && ends && feat && grad_feat, "rsn_composite16_bwd: null pointer");
  RSN_ARG(!grad_rgb_blend || (feat_out && accumulation), "rsn_composite16_bwd: feat_out/accumulation required with grad_rgb_blend");
  return launch_bwd<16>(sigma, starts, ends, bin_row_stride, feat, grad_weights, grad_accumulation, grad_feat_out,
                        grad_sigma, grad_feat, n_rays, (int)n_samples, stream, n_rays_dev, normals, grad_pred_normal_loss,
                        grad_orientation_loss, grad_rgb_blend, feat_out, accumulation);
}


extern "C" int rsn_render_weights(const float* weights, const float* feat, int64_t n_channels, const float* starts,
                                  const float* ends, int64_t bin_row_stride, float* accumulation, float* feat_out,
                                  float* depth_median, int64_t n_rays, int64_t n_samples, cudaStream_t stream) {
  RSN_ARG(n_rays >= 0 && n_samples >= 1 && n_channels >= 0 && n_channels <= 16, "rsn_render_weights: bad shape");
  if (n_rays == 0) return 0;
  RSN_ARG(weights != nullptr, "rsn_render_weights: null pointer");
  RSN_ARG(n_channels == 0 || (feat && feat_out), "rsn_render_weights: feat/feat_out required when n_channels > 0");
  RSN_ARG(!depth_median || (starts && ends), "rsn_render_weights: starts/ends required for the median depth");
  render_weights_kernel<<<(unsigned)((n_rays * 32 + 255) / 256), 256, 0, stream>>>(
      weights, feat, (int)n_channels, starts, ends, bin_row_stride, accumulation, feat_out, depth_median, n_rays,
      (int)n_samples);
  RSN_LAUNCH_CHECK("render_weights_kernel");
  return 0;
}
