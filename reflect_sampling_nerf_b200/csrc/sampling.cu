// K1 (spaced sampling) and K2 (PDF resampling) -- SURVEY.md §2.4, §8 rows a2-a4.
//
// Replaces (reference call sites; arithmetic restated in oracle/upstream.py):
//   UniformSampler / ReciprocalSampler  reflect_sampling_nerf_model.py:109,111,148,292
//                                        reflect_sampling_nerf_components.py:14-36
//   PDFSampler(include_original=False)   reflect_sampling_nerf_model.py:110,112,182,317
//
// These kernels are "bit-exact" targets: every fp32 operation is issued with an explicit
// round-to-nearest intrinsic so nvcc cannot contract a*b+c into an FMA, and the operation order is
// the oracle's.  The per-ray sum and the CDF cumsum accumulate in fp64 and round once per output
// (oracle/upstream.py, SUM_MODE == "fp64"; torch's CPU cumsum already behaves that way).
#include "rsn_common.cuh"
#include <algorithm>

#ifndef RSN_PDF_MINB
#define RSN_PDF_MINB 1
#endif
namespace {

// ReciprocalSampler (reflect_sampling_nerf_components.py:30-33): spacing_fn = x / (1/tan + x), inverse = x / tan / (1 - x);
// the python scalars 1/tan and tan reach the fp32 tensor ops rounded to fp32 (tan = 0.25 in the model: 4.0 and 0.25 exactly)
struct Spacing {
  int kind;        // 0 uniform (identity), 1 reciprocal
  float inv_tan;   // fl32(1 / tan)
  float tan;       // fl32(tan)
};
__device__ __forceinline__ float spacing_fn(float x, Spacing sp) {
  return sp.kind == 0 ? x : __fdiv_rn(x, __fadd_rn(sp.inv_tan, x));
}
__device__ __forceinline__ float spacing_inv(float x, Spacing sp) {
  return sp.kind == 0 ? x : __fdiv_rn(__fdiv_rn(x, sp.tan), __fsub_rn(1.0f, x));
}
__device__ __forceinline__ float to_euclid(float b, float s_near, float s_far, Spacing sp) {
  // x * s_far + (1 - x) * s_near, each op rounded separately
  float v = __fadd_rn(__fmul_rn(b, s_far), __fmul_rn(__fsub_rn(1.0f, b), s_near));
  return spacing_inv(v, sp);
}

// One thread per bin.  HBM-bound: 8 B written per bin (+4 B read when jitter is injected).  IDX = unsigned (32-bit index
// arithmetic, whenever the pass has fewer than 2^31 bins) or int64_t: the 64-bit division of the flat index by the
// run-time bin count was most of the kernel's instructions (ncu: 68 % issue-active at 13 % of the DRAM peak).  x / 2 is
// written as x * 0.5 (identical in IEEE arithmetic, one instruction instead of a division routine).
template <class IDX>
__global__ void __launch_bounds__(256) sample_spaced_kernel(
    const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ lin,
    const float* __restrict__ t_rand, int64_t t_rand_cols, Spacing kind, float* __restrict__ spacing,
    float* __restrict__ euclid, int64_t n_rays, int n_bins, const int* __restrict__ n_rays_dev) {
  IDX idx = (IDX)blockIdx.x * blockDim.x + threadIdx.x;
  const IDX total = (IDX)(rsn_count(n_rays, n_rays_dev) * n_bins), step = (IDX)gridDim.x * blockDim.x;
  const IDX nb = (IDX)n_bins;
  for (; idx < total; idx += step) {
    const IDX r = idx / nb;
    const int i = (int)(idx - r * nb);
    float b = __ldg(lin + i);
    if (t_rand != nullptr) {
      float t = __ldg(t_rand + (int64_t)r * t_rand_cols + (t_rand_cols == 1 ? 0 : i));
      float lower = i == 0 ? b : __fmul_rn(__fadd_rn(b, __ldg(lin + i - 1)), 0.5f);
      float upper = i == n_bins - 1 ? b : __fmul_rn(__fadd_rn(__ldg(lin + i + 1), b), 0.5f);
      b = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t));
    }
    float s_near = spacing_fn(__ldg(nears + r), kind);
    float s_far = spacing_fn(__ldg(fars + r), kind);
    spacing[idx] = b;
    euclid[idx] = to_euclid(b, s_near, s_far, kind);
  }
}

// The same per (bin, ray) arithmetic with one THREAD PER BIN COLUMN walking `rays_per_block` rays: everything that depends
// on the bin only (lin[i], the stratification interval) is computed once per thread, no index division at all, and a
// warp's accesses to a row are contiguous.  Used whenever a row fits a block (n_bins <= 1024).
__global__ void __launch_bounds__(1024) sample_spaced_rows_kernel(
    const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ lin,
    const float* __restrict__ t_rand, int64_t t_rand_cols, Spacing kind, float* __restrict__ spacing,
    float* __restrict__ euclid, int64_t n_rays, int n_bins, int rays_per_block, const int* __restrict__ n_rays_dev) {
  const int i = threadIdx.x;
  if (i >= n_bins) return;
  const int64_t n_valid = rsn_count(n_rays, n_rays_dev);
  const int64_t r0 = (int64_t)blockIdx.x * rays_per_block;
  const int64_t r1 = r0 + rays_per_block < n_valid ? r0 + rays_per_block : n_valid;
  const float b0 = __ldg(lin + i);
  float lower = b0, width = 0.f;
  if (t_rand != nullptr) {
    lower = i == 0 ? b0 : __fmul_rn(__fadd_rn(b0, __ldg(lin + i - 1)), 0.5f);
    const float upper = i == n_bins - 1 ? b0 : __fmul_rn(__fadd_rn(__ldg(lin + i + 1), b0), 0.5f);
    width = __fsub_rn(upper, lower);
  }
  const int64_t tcol = t_rand_cols == 1 ? 0 : i;
#pragma unroll 4
  for (int64_t r = r0; r < r1; ++r) {
    float b = b0;
    if (t_rand != nullptr) b = __fadd_rn(lower, __fmul_rn(width, __ldg(t_rand + r * t_rand_cols + tcol)));
    const float s_near = spacing_fn(__ldg(nears + r), kind);
    const float s_far = spacing_fn(__ldg(fars + r), kind);
    spacing[r * n_bins + i] = b;
    euclid[r * n_bins + i] = to_euclid(b, s_near, s_far, kind);
  }
}

// The flat form: the [N, n_bins] arrays are contiguous, so a thread takes FOUR consecutive flat elements -- one 128-bit
// load of the noise, two 128-bit stores -- and splits the flat index into (ray, bin) once (32-bit division), stepping the
// pair for the other three.  Same per-element arithmetic.  Needs 16-byte aligned arrays and fewer than 2^31 bins per pass.
__global__ void __launch_bounds__(256) sample_spaced_vec4_kernel(
    const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ lin,
    const float* __restrict__ t_rand, Spacing kind, float* __restrict__ spacing, float* __restrict__ euclid, int64_t n_rays,
    int n_bins, const int* __restrict__ n_rays_dev) {
  const unsigned total = (unsigned)(rsn_count(n_rays, n_rays_dev) * n_bins);
  const unsigned nb = (unsigned)n_bins;
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q * 4u < total; q += gridDim.x * blockDim.x) {
    const unsigned idx = q * 4u;
    unsigned r = idx / nb, i = idx - r * nb;
    float t[4] = {0.f, 0.f, 0.f, 0.f}, sp[4], eu[4];
    const bool full = idx + 4u <= total;
    if (t_rand != nullptr) {
      if (full) {
        const float4 tv = __ldg(reinterpret_cast<const float4*>(t_rand + idx));
        t[0] = tv.x, t[1] = tv.y, t[2] = tv.z, t[3] = tv.w;
      } else {
        for (unsigned k = 0; idx + k < total; ++k) t[k] = __ldg(t_rand + idx + k);
      }
    }
    float s_near = spacing_fn(__ldg(nears + r), kind), s_far = spacing_fn(__ldg(fars + r), kind);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float b = __ldg(lin + i);
      if (t_rand != nullptr) {
        const float lower = i == 0 ? b : __fmul_rn(__fadd_rn(b, __ldg(lin + i - 1)), 0.5f);
        const float upper = i == nb - 1 ? b : __fmul_rn(__fadd_rn(__ldg(lin + i + 1), b), 0.5f);
        b = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t[k]));
      }
      sp[k] = b;
      eu[k] = to_euclid(b, s_near, s_far, kind);
      if (++i == nb && k < 3) {   // next ray
        i = 0;
        ++r;
        if (idx + k + 1 < total) s_near = spacing_fn(__ldg(nears + r), kind), s_far = spacing_fn(__ldg(fars + r), kind);
      }
    }
    if (full) {
      *reinterpret_cast<float4*>(spacing + idx) = make_float4(sp[0], sp[1], sp[2], sp[3]);
      *reinterpret_cast<float4*>(euclid + idx) = make_float4(eu[0], eu[1], eu[2], eu[3]);
    } else {
      for (unsigned k = 0; idx + k < total; ++k) spacing[idx + k] = sp[k], euclid[idx + k] = eu[k];
    }
  }
}

// One warp per ray.  smem per warp: cdf[S+1], the existing spacing bins[S+1] and the ray's weights (+ histogram padding),
// staged with coalesced loads (lane-strided) and read back by the lane that owns the sample's chunk; the staging row is
// skewed by one word per 32 so that the chunked reads (lane stride = chunk words) hit distinct banks.  The arithmetic
// and its order are unchanged (fp64 partial sums per lane-chunk, xor-tree, fp64 scan), so the bins stay bit-exact.
// searchsorted(cdf[0..S], u, side="right") = number of entries <= u, as a fixed-trip-count, branch-free descent (the CDF is
// non-decreasing: a cumulative sum of non-negative terms, clamped at 1) -- the same index the bisection loop finds
template <int MAXSTEP>
__device__ __forceinline__ int count_le(const float* cdf, int n, float u) {
  int pos = 0;
#pragma unroll
  for (int step = MAXSTEP; step >= 1; step >>= 1) {
    const int nxt = pos + step;
    if (nxt <= n && cdf[nxt - 1] <= u) pos = nxt;
  }
  return pos;
}

// CHUNK = ceil(S / 32) known at compile time (S <= 256): every per-sample loop is unrolled and predicated, the PDF value of
// a sample is divided once and kept in a register between the sum and the CDF pass, the search is branch-free.  Same
// arithmetic in the same order as the generic kernel below (ncu of the generic form: 1,370 warp instructions per ray, 68 %
// issue-active, two thirds of them loop and address bookkeeping).
// S_T > 0: additionally specialised on the sample count itself (the model's S = 128 and 64 with S + 1 output bins): every
// predicate of the unrolled loops folds away.
template <int WARPS, int CHUNK, int S_T>
__global__ void __launch_bounds__(WARPS * 32, RSN_PDF_MINB) pdf_resample_chunk_kernel(
    const float* __restrict__ weights, int64_t w_stride, const float* __restrict__ bins_in,
    const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ u_base,
    const float* __restrict__ rand, Spacing kind, float hist_pad, float* __restrict__ spacing_out,
    float* __restrict__ euclid_out, int64_t* __restrict__ inds_out, int64_t n_rays, int S_rt, int nb_rt,
    const int* __restrict__ n_rays_dev) {
  const int S = S_T ? S_T : S_rt, nb = S_T ? S_T + 1 : nb_rt;
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int OUT = CHUNK + 1;    // consecutive outputs per lane in the fast output phase (needs nb <= 32 OUT)
  const bool fast_out = nb <= 32 * OUT;
  const int row_words = 2 * (S + 1) + S + (S >> 5) + 1 + (fast_out ? 3 * nb : 0);
  float* cdf = smem + (size_t)warp * row_words;
  float* ebins = cdf + (S + 1);
  float* wp = ebins + (S + 1);      // wp[i + (i >> 5)] = weights[i] + histogram_padding
  float* rnd_s = wp + S + (S >> 5) + 1;   // fast output phase: the ray's noise, then its two output rows
  float* out_s = rnd_s + nb;
  const int chunk = (S + 31) / 32;  // consecutive samples per lane (<= CHUNK)
  const int64_t n_valid = rsn_count(n_rays, n_rays_dev);
  constexpr int MAXSTEP = CHUNK <= 1 ? 32 : CHUNK <= 3 ? 64 : CHUNK <= 7 ? 128 : 256;   // largest power of two <= 32 CHUNK + 1
  // u_base of this lane's outputs (the same for every ray)
  float ub[OUT];
#pragma unroll
  for (int k = 0; k < OUT; ++k) ub[k] = (fast_out && lane * OUT + k < nb) ? __ldg(u_base + lane * OUT + k) : 0.f;

  for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < n_valid; r += (int64_t)gridDim.x * WARPS) {
    const float* w = weights + r * w_stride;
    const int lo = lane * chunk;
    float rr[OUT];   // the ray's noise, loaded coalesced now and staged for the per-lane consecutive outputs after the CDF
    if (fast_out && rand != nullptr) {
#pragma unroll
      for (int k = 0; k < OUT; ++k) rr[k] = (lane + 32 * k < nb) ? __ldg(rand + r * nb + lane + 32 * k) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < CHUNK; ++it) {
      const int i = lane + 32 * it;
      if (i < S) wp[i + (i >> 5)] = __fadd_rn(__ldg(w + i), hist_pad);
    }
#pragma unroll
    for (int it = 0; it <= CHUNK; ++it) {
      const int i = lane + 32 * it;
      if (i <= S) ebins[i] = __ldg(bins_in + r * (S + 1) + i);
    }
    __syncwarp();
    float wv[CHUNK];
    double part = 0.0;
#pragma unroll
    for (int k = 0; k < CHUNK; ++k) {
      const int i = lo + k;
      const bool ok = k < chunk && i < S;
      wv[k] = ok ? wp[i + (i >> 5)] : 0.f;
      if (ok) part += (double)wv[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(RSN_FULL, part, o);
    float wsum = (float)part;
    const float padding = fmaxf(__fsub_rn(1e-5f, wsum), 0.0f);  // relu(eps - sum)
    const float pad_each = __fdiv_rn(padding, (float)S);
    wsum = __fadd_rn(wsum, padding);
    double run = 0.0;
#pragma unroll
    for (int k = 0; k < CHUNK; ++k) {
      const bool ok = k < chunk && lo + k < S;
      wv[k] = ok ? __fdiv_rn(__fadd_rn(wv[k], pad_each), wsum) : 0.f;    // the PDF value, divided once
      if (ok) run += (double)wv[k];
    }
    const double incl = warp_incl_scan(run, lane);
    double acc = incl - run;  // exclusive offset of this lane's chunk
#pragma unroll
    for (int k = 0; k < CHUNK; ++k) {
      const int i = lo + k;
      if (k < chunk && i < S) {
        acc += (double)wv[k];
        cdf[i + 1] = fminf(1.0f, (float)acc);
      }
    }
    if (lane == 0) cdf[0] = 0.0f;
    if (fast_out && rand != nullptr) {
#pragma unroll
      for (int k = 0; k < OUT; ++k)
        if (lane + 32 * k < nb) rnd_s[lane + 32 * k] = rr[k];
    }
    __syncwarp();

    const float s_near = spacing_fn(__ldg(nears + r), kind);
    const float s_far = spacing_fn(__ldg(fars + r), kind);
    const float nbf = (float)nb;
    if (fast_out) {
      // Lane l owns the OUT consecutive outputs j = l OUT + k.  Their u are (almost always) non-decreasing, so only the first
      // index is searched; the following ones advance from their predecessor (a fresh search if a u steps back by an ulp).
      int a = 0;
      float u_prev = 0.f;
#pragma unroll
      for (int k = 0; k < OUT; ++k) {
        const int j = lane * OUT + k;
        if (j < nb) {
          float u = ub[k];
          if (rand != nullptr) u = __fadd_rn(u, __fdiv_rn(rnd_s[j], nbf));
          if (k == 0 || u < u_prev) {
            a = count_le<MAXSTEP>(cdf, S + 1, u);
          } else {
            while (a <= S && cdf[a] <= u) ++a;
          }
          u_prev = u;
          const int below = min(max(a - 1, 0), S), above = min(a, S);
          const float c0 = cdf[below], c1 = cdf[above], b0 = ebins[below], b1 = ebins[above];
          float t = nan_to_num(__fdiv_rn(__fsub_rn(u, c0), __fsub_rn(c1, c0)));
          t = fminf(fmaxf(t, 0.0f), 1.0f);
          const float nbv = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
          out_s[j] = nbv;
          out_s[nb + j] = to_euclid(nbv, s_near, s_far, kind);
          if (inds_out != nullptr) inds_out[r * nb + j] = a;
        }
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < OUT; ++k) {   // coalesced write-out
        const int j = lane + 32 * k;
        if (j < nb) {
          spacing_out[r * nb + j] = out_s[j];
          euclid_out[r * nb + j] = out_s[nb + j];
        }
      }
      __syncwarp();
      continue;
    }
    for (int j = lane; j < nb; j += 32) {
      float u = __ldg(u_base + j);
      if (rand != nullptr) u = __fadd_rn(u, __fdiv_rn(__ldg(rand + r * nb + j), nbf));
      const int a = count_le<MAXSTEP>(cdf, S + 1, u);
      const int below = min(max(a - 1, 0), S), above = min(a, S);
      const float c0 = cdf[below], c1 = cdf[above], b0 = ebins[below], b1 = ebins[above];
      float t = nan_to_num(__fdiv_rn(__fsub_rn(u, c0), __fsub_rn(c1, c0)));
      t = fminf(fmaxf(t, 0.0f), 1.0f);
      const float nbv = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
      spacing_out[r * nb + j] = nbv;
      euclid_out[r * nb + j] = to_euclid(nbv, s_near, s_far, kind);
      if (inds_out != nullptr) inds_out[r * nb + j] = a;
    }
    __syncwarp();
  }
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) pdf_resample_kernel(
    const float* __restrict__ weights, int64_t w_stride, const float* __restrict__ bins_in,
    const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ u_base,
    const float* __restrict__ rand, Spacing kind, float hist_pad, float* __restrict__ spacing_out,
    float* __restrict__ euclid_out, int64_t* __restrict__ inds_out, int64_t n_rays, int S, int nb,
    const int* __restrict__ n_rays_dev) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_words = 2 * (S + 1) + S + (S >> 5) + 1;
  float* cdf = smem + (size_t)warp * row_words;
  float* ebins = cdf + (S + 1);
  float* wp = ebins + (S + 1);      // wp[i + (i >> 5)] = weights[i] + histogram_padding
  const int chunk = (S + 31) / 32;  // consecutive samples per lane
  const int64_t n_valid = rsn_count(n_rays, n_rays_dev);

  for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < n_valid; r += (int64_t)gridDim.x * WARPS) {
    const float* w = weights + r * w_stride;
    const int lo = lane * chunk, hi = min(S, lo + chunk);
    for (int i = lane; i < S; i += 32) wp[i + (i >> 5)] = __fadd_rn(__ldg(w + i), hist_pad);
    for (int i = lane; i <= S; i += 32) ebins[i] = __ldg(bins_in + r * (S + 1) + i);
    __syncwarp();
    // weights + histogram_padding, row sum in fp64
    double part = 0.0;
    for (int i = lo; i < hi; ++i) part += (double)wp[i + (i >> 5)];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(RSN_FULL, part, o);
    float wsum = (float)part;
    float padding = fmaxf(__fsub_rn(1e-5f, wsum), 0.0f);  // relu(eps - sum)
    float pad_each = __fdiv_rn(padding, (float)S);
    wsum = __fadd_rn(wsum, padding);
    // pdf and its inclusive cumsum (fp64 accumulate, rounded per element), clamped at 1
    double run = 0.0;
    for (int i = lo; i < hi; ++i) {
      float wi = __fadd_rn(wp[i + (i >> 5)], pad_each);
      run += (double)__fdiv_rn(wi, wsum);
    }
    double incl = warp_incl_scan(run, lane);
    double acc = incl - run;  // exclusive offset of this lane's chunk
    for (int i = lo; i < hi; ++i) {
      float wi = __fadd_rn(wp[i + (i >> 5)], pad_each);
      acc += (double)__fdiv_rn(wi, wsum);
      cdf[i + 1] = fminf(1.0f, (float)acc);
    }
    if (lane == 0) cdf[0] = 0.0f;
    __syncwarp();

    const float s_near = spacing_fn(__ldg(nears + r), kind);
    const float s_far = spacing_fn(__ldg(fars + r), kind);
    for (int j = lane; j < nb; j += 32) {
      float u = __ldg(u_base + j);
      if (rand != nullptr) u = __fadd_rn(u, __fdiv_rn(__ldg(rand + r * nb + j), (float)nb));
      // searchsorted(cdf, u, side="right"): number of entries <= u
      int a = 0, b = S + 1;
      while (a < b) {
        int m = (a + b) >> 1;
        if (cdf[m] <= u) a = m + 1; else b = m;
      }
      const int below = min(max(a - 1, 0), S), above = min(max(a, 0), S);
      const float c0 = cdf[below], c1 = cdf[above], b0 = ebins[below], b1 = ebins[above];
      float t = nan_to_num(__fdiv_rn(__fsub_rn(u, c0), __fsub_rn(c1, c0)));
      t = fminf(fmaxf(t, 0.0f), 1.0f);
      const float nbv = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
      spacing_out[r * nb + j] = nbv;
      euclid_out[r * nb + j] = to_euclid(nbv, s_near, s_far, kind);
      if (inds_out != nullptr) inds_out[r * nb + j] = a;
    }
    __syncwarp();
  }
}

}  // namespace

extern "C" int rsn_sample_spaced(const float* nears, const float* fars, const float* lin_bins,
                                 const float* t_rand, int64_t t_rand_cols, int spacing_kind, float tan,
                                 float* spacing_bins, float* euclid_bins, int64_t n_rays,
                                 int64_t n_samples, const int* n_rays_dev, cudaStream_t stream) {
  RSN_ARG(n_rays >= 0 && n_samples >= 1, "rsn_sample_spaced: bad shape (%lld rays, %lld samples)",
          (long long)n_rays, (long long)n_samples);
  RSN_ARG(spacing_kind == 0 || (spacing_kind == 1 && tan > 0.f), "rsn_sample_spaced: spacing_kind must be 0 (uniform) or 1 (reciprocal, tan > 0)");
  const Spacing sp = {spacing_kind, (float)(1.0 / (double)tan), tan};
  RSN_ARG(t_rand == nullptr || t_rand_cols == 1 || t_rand_cols == n_samples + 1,
          "rsn_sample_spaced: t_rand must have 1 or n_samples+1 columns");
  if (n_rays == 0) return 0;
  RSN_ARG(nears && fars && lin_bins && spacing_bins && euclid_bins, "rsn_sample_spaced: null pointer");
  const int n_bins = (int)n_samples + 1;
  int64_t total = n_rays * n_bins;
  int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)rsn_num_sms() * 8);
  const bool vec_ok = total < (int64_t)2147483647 - 4096 && (t_rand == nullptr || t_rand_cols == n_bins) &&
                      (((uintptr_t)t_rand | (uintptr_t)spacing_bins | (uintptr_t)euclid_bins) & 15) == 0;
  if (vec_ok) {
    const int64_t quads = (total + 3) / 4;
    sample_spaced_vec4_kernel<<<(unsigned)std::min<int64_t>((quads + 255) / 256, (int64_t)rsn_num_sms() * 16), 256, 0, stream>>>(
        nears, fars, lin_bins, t_rand, sp, spacing_bins, euclid_bins, n_rays, n_bins, n_rays_dev);
  } else if (n_bins <= 1024) {
    const int threads = (n_bins + 31) & ~31;
    const int rpb = 16;
    sample_spaced_rows_kernel<<<(unsigned)((n_rays + rpb - 1) / rpb), threads, 0, stream>>>(
        nears, fars, lin_bins, t_rand, t_rand_cols, sp, spacing_bins, euclid_bins, n_rays, n_bins, rpb, n_rays_dev);
  } else if (total < (int64_t)2147483647 - (int64_t)blocks * 256)
    sample_spaced_kernel<unsigned><<<blocks, 256, 0, stream>>>(nears, fars, lin_bins, t_rand, t_rand_cols, sp, spacing_bins,
                                                               euclid_bins, n_rays, n_bins, n_rays_dev);
  else
    sample_spaced_kernel<int64_t><<<blocks, 256, 0, stream>>>(nears, fars, lin_bins, t_rand, t_rand_cols, sp, spacing_bins,
                                                              euclid_bins, n_rays, n_bins, n_rays_dev);
  RSN_LAUNCH_CHECK("sample_spaced_kernel");
  return 0;
}

extern "C" int rsn_pdf_resample(const float* weights, int64_t weights_row_stride, const float* spacing_bins_in,
                                const float* nears, const float* fars, const float* u_base, const float* rand,
                                int spacing_kind, float tan, float histogram_padding, float* spacing_bins_out,
                                float* euclid_bins_out, int64_t* inds_out, int64_t n_rays, int64_t n_in_samples,
                                int64_t n_out_samples, const int* n_rays_dev, cudaStream_t stream) {
  RSN_ARG(n_rays >= 0 && n_in_samples >= 1 && n_out_samples >= 1, "rsn_pdf_resample: bad shape");
  RSN_ARG(n_in_samples <= 4096, "rsn_pdf_resample: at most 4096 input samples per ray (got %lld)", (long long)n_in_samples);
  RSN_ARG(spacing_kind == 0 || (spacing_kind == 1 && tan > 0.f), "rsn_pdf_resample: spacing_kind must be 0 or 1 (tan > 0)");
  const Spacing sp = {spacing_kind, (float)(1.0 / (double)tan), tan};
  if (n_rays == 0) return 0;
  RSN_ARG(weights && spacing_bins_in && nears && fars && u_base && spacing_bins_out && euclid_bins_out,
          "rsn_pdf_resample: null pointer");
  constexpr int WARPS = 4;
  const int S = (int)n_in_samples, nb = (int)n_out_samples + 1;
  // (+ 3 nb floats per warp when the specialised kernel's fast output phase applies: nb <= 32 (ceil(S / 32) + 1))
  const int chunk_t = (S + 31) / 32 <= 1 ? 1 : (S + 31) / 32 <= 2 ? 2 : (S + 31) / 32 <= 4 ? 4 : 8;
  const bool fast_out = (S + 31) / 32 <= 8 && nb <= 32 * (chunk_t + 1);
  size_t smem = (size_t)WARPS * (2 * (S + 1) + S + (S >> 5) + 1 + (fast_out ? 3 * nb : 0)) * sizeof(float);
  static std::atomic<unsigned long long> done;   // sized for the 4096-sample maximum, set once per device
  RSN_CUDA(rsn_ensure_smem(pdf_resample_kernel<WARPS>, (int)(WARPS * (2 * 4097 + 4096 + 129) * sizeof(float)), done));
  int blocks = (int)std::min<int64_t>((n_rays + WARPS - 1) / WARPS, (int64_t)rsn_num_sms() * 16);
#define RSN_PDF_CHUNK(C, ST)                                                                                  \
  pdf_resample_chunk_kernel<WARPS, C, ST><<<blocks, WARPS * 32, smem, stream>>>(                              \
      weights, weights_row_stride, spacing_bins_in, nears, fars, u_base, rand, sp, histogram_padding,        \
      spacing_bins_out, euclid_bins_out, inds_out, n_rays, S, nb, n_rays_dev)
  const int chunk = (S + 31) / 32;
  if (S == 128 && nb == 129) RSN_PDF_CHUNK(4, 128);       // the model's primary passes
  else if (S == 64 && nb == 65) RSN_PDF_CHUNK(2, 64);     // ... and its bounce passes
  else if (chunk <= 1) RSN_PDF_CHUNK(1, 0);
  else if (chunk <= 2) RSN_PDF_CHUNK(2, 0);
  else if (chunk <= 4) RSN_PDF_CHUNK(4, 0);
  else if (chunk <= 8) RSN_PDF_CHUNK(8, 0);
  else
    pdf_resample_kernel<WARPS><<<blocks, WARPS * 32, smem, stream>>>(
        weights, weights_row_stride, spacing_bins_in, nears, fars, u_base, rand, sp, histogram_padding,
        spacing_bins_out, euclid_bins_out, inds_out, n_rays, S, nb, n_rays_dev);
#undef RSN_PDF_CHUNK
  RSN_LAUNCH_CHECK("pdf_resample_kernel");
  return 0;
}
