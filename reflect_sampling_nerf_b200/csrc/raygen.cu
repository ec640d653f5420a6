// Pixel sampling + camera ray generation + target-pixel gather in one launch -- SURVEY.md §8 row f1.
//
// Replaces, per training batch (reflect_sampling_nerf_datamanager.py:49-58 -> upstream PixelSampler.sample_method,
// RayGenerator -> Cameras.generate_rays / _generate_rays_from_coords, perspective cameras without distortion, and the
// Blender parser's alpha blend onto white):
//   (cam, y, x) = floor(rand[N,3] * (V, H, W))                                       PixelSampler
//   coord = ((x + .5 - cx) / fx, -(y + .5 - cy) / fy, -1), its x+1 / y+1 neighbours   Cameras._generate_rays_from_coords
//   d = normalize(R coord) for the three, origin = c2w[:, 3]
//   pixel_area = |d - d_x| |d - d_y|
//   target = rgb / 255 * a / 255 + (1 - a / 255)                                      blender_dataparser (alpha_color = white)
// At > 500 k rays/s the CPU dataloader -> RayGenerator hop of the reference is the bottleneck of the real training loop;
// here the images (uint8 RGBA) and cameras stay resident in HBM and a batch costs one ~10 us launch.
// Every fp32 operation is an explicit round-to-nearest intrinsic in the oracle's order (oracle/cameras.py): bit-exact.
// HBM: 12 B (rand) + 4 B (pixel) in, 44 B out per ray.
#include "rsn_common.cuh"

namespace {

__global__ void __launch_bounds__(256) raygen_kernel(const float* __restrict__ c2w, const float* __restrict__ intr,
                                                     const float* __restrict__ rnd, const int64_t* __restrict__ pix_in,
                                                     const uint8_t* __restrict__ images, int n_channels, int V, int H, int W,
                                                     float* __restrict__ origins, float* __restrict__ dirs,
                                                     float* __restrict__ area, int64_t* __restrict__ pix_out,
                                                     float* __restrict__ target, int64_t n) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int cam, yi, xi;
  if (pix_in) {
    cam = (int)__ldg(pix_in + r * 3), yi = (int)__ldg(pix_in + r * 3 + 1), xi = (int)__ldg(pix_in + r * 3 + 2);
  } else {   // floor(rand * (V, H, W)).long(); rand < 1, the product can still round up to the bound
    cam = min((int)floorf(__fmul_rn(__ldg(rnd + r * 3), (float)V)), V - 1);
    yi = min((int)floorf(__fmul_rn(__ldg(rnd + r * 3 + 1), (float)H)), H - 1);
    xi = min((int)floorf(__fmul_rn(__ldg(rnd + r * 3 + 2), (float)W)), W - 1);
  }
  const float fx = __ldg(intr + cam * 4), fy = __ldg(intr + cam * 4 + 1), cx = __ldg(intr + cam * 4 + 2),
              cy = __ldg(intr + cam * 4 + 3);
  const float x = __fadd_rn((float)xi, 0.5f), y = __fadd_rn((float)yi, 0.5f);   // image coordinates at pixel centres
  const float xc = __fsub_rn(x, cx), yc = __fsub_rn(y, cy);
  const float u[3] = {__fdiv_rn(xc, fx), __fdiv_rn(__fadd_rn(xc, 1.0f), fx), __fdiv_rn(xc, fx)};
  const float v[3] = {-__fdiv_rn(yc, fy), -__fdiv_rn(yc, fy), -__fdiv_rn(__fadd_rn(yc, 1.0f), fy)};
  const float* m = c2w + (size_t)cam * 12;   // [3][4] row-major
  float R[3][3], t[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) R[i][j] = __ldg(m + i * 4 + j);
    t[i] = __ldg(m + i * 4 + 3);
  }
  float d[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float w[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)   // sum_j coord_j R_ij, coord = (u, v, -1)
      w[i] = __fadd_rn(__fadd_rn(__fmul_rn(u[k], R[i][0]), __fmul_rn(v[k], R[i][1])), __fmul_rn(-1.0f, R[i][2]));
    const float nrm = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(w[0], w[0]), __fmul_rn(w[1], w[1])), __fmul_rn(w[2], w[2]))),
                            8.881784197001252e-16f);   // camera_utils._EPS = 4 * float64 eps
#pragma unroll
    for (int i = 0; i < 3; ++i) d[k][i] = __fdiv_rn(w[i], nrm);
  }
  float dxy[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float a = __fsub_rn(d[0][0], d[k + 1][0]), b = __fsub_rn(d[0][1], d[k + 1][1]), c = __fsub_rn(d[0][2], d[k + 1][2]);
    dxy[k] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)));
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    origins[r * 3 + i] = t[i];
    dirs[r * 3 + i] = d[0][i];
  }
  area[r] = __fmul_rn(dxy[0], dxy[1]);
  if (pix_out) pix_out[r * 3] = cam, pix_out[r * 3 + 1] = yi, pix_out[r * 3 + 2] = xi;
  if (images && target) {
    const uint8_t* px = images + (((size_t)cam * H + yi) * W + xi) * n_channels;
    float c[3] = {__fdiv_rn((float)px[0], 255.0f), __fdiv_rn((float)px[1], 255.0f), __fdiv_rn((float)px[2], 255.0f)};
    if (n_channels == 4) {
      const float a = __fdiv_rn((float)px[3], 255.0f), rest = __fsub_rn(1.0f, a);
#pragma unroll
      for (int i = 0; i < 3; ++i) c[i] = __fadd_rn(__fmul_rn(c[i], a), rest);
    }
    target[r * 3] = c[0], target[r * 3 + 1] = c[1], target[r * 3 + 2] = c[2];
  }
}

}  // namespace

extern "C" int rsn_raygen(const float* c2w, const float* intrinsics, const float* rand3, const int64_t* pixels_in,
                          const uint8_t* images, int64_t n_channels, int64_t n_views, int64_t height, int64_t width,
                          float* origins, float* dirs, float* pixel_area, int64_t* pixels_out, float* target_rgb,
                          int64_t n_rays, cudaStream_t stream) {
  RSN_ARG(n_rays >= 0 && n_views >= 1 && height >= 1 && width >= 1, "rsn_raygen: bad shape");
  if (n_rays == 0) return 0;
  RSN_ARG(c2w && intrinsics && origins && dirs && pixel_area, "rsn_raygen: null pointer");
  RSN_ARG((rand3 != nullptr) != (pixels_in != nullptr), "rsn_raygen: exactly one of rand3 / pixels_in");
  RSN_ARG(!images || (target_rgb && (n_channels == 3 || n_channels == 4)), "rsn_raygen: images need target_rgb and 3 or 4 channels");
  raygen_kernel<<<(unsigned)((n_rays + 255) / 256), 256, 0, stream>>>(c2w, intrinsics, rand3, pixels_in, images, (int)n_channels,
                                                                      (int)n_views, (int)height, (int)width, origins, dirs,
                                                                      pixel_area, pixels_out, target_rgb, n_rays);
  RSN_LAUNCH_CHECK("raygen_kernel");
  return 0;
}
