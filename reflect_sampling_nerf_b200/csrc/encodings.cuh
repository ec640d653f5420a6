// Encodings shared by the fused field kernel (csrc/field_fwd.cu) and the stand-alone encoders of the component API
// (csrc/encode.cu): the integrated positional encoding of NeRFEncoding.forward(x, covs) (SURVEY.md App. A.4;
// reflect_sampling_nerf_model.py:98-100) and IntegratedSHEncoding (reflect_sampling_nerf_components.py:52-140).
#pragma once
#include "rsn_common.cuh"

namespace rsnenc {

// 2 ** torch.linspace(0, 16, 16) in fp32, bit for bit (NeRFEncoding, reflect_sampling_nerf_model.py:98-100)
static __constant__ float c_freq[16] = {
    0x1.0000000000000p+0f,  0x1.0c1b780000000p+1f,  0x1.18c9880000000p+2f,  0x1.26111c0000000p+3f,
    0x1.33f9760000000p+4f,  0x1.428a320000000p+5f,  0x1.51cb4e0000000p+6f,  0x1.61c5140000000p+7f,
    0x1.7280340000000p+8f,  0x1.8405f60000000p+9f,  0x1.965fde0000000p+10f, 0x1.a998080000000p+11f,
    0x1.bdb8d20000000p+12f, 0x1.d2cd4c0000000p+13f, 0x1.e8e1020000000p+14f, 0x1.0000000000000p+16f};

// sin of an fp32 argument of any magnitude the encoding produces (|s| <= 2pi * 2 * 65536): two-term
// Cody-Waite reduction by 2pi (exact products through fma), then the SFU on [-pi, pi].  Absolute error
// < 1e-6, far below the bf16 resolution of the feature it feeds.
__device__ __forceinline__ float sin_reduced(float s) {
  const float k = rintf(s * 0.15915494309189535f);
  float r = fmaf(-k, 0x1.921fb60000000p+2f, s);
  r = fmaf(-k, -0x1.777a5cp-23f, r);
  return __sinf(r);
}

// One IPE feature: exp(-0.5 var f^2) sin(2 pi x f [+ pi/2]); sx = fl(2 pi x), va = diag(cov) entry, f = c_freq[k].
// s = fl(fl(2 pi x) f) follows the reference's fp32 operation order (it feeds sin at f up to 65536); features damped
// below e^-24 are exact zeros.
__device__ __forceinline__ float ipe_value(float sx, float va, float f, int half) {
  float s = __fmul_rn(sx, f);
  if (half) s = __fadd_rn(s, 1.5707963705062866f);
  const float e = 0.5f * (va * (f * f));
  return (e < 24.f) ? __expf(-e) * sin_reduced(s) : 0.f;
}

// ---------------------------------------------------------------------------------------------- IDE
// IntegratedSHEncoding.pytorch_fwd (components.py:52-140): 34 hand-expanded polynomials for l = 1,2,4,8 with
// the reference's constants (entries 17/18/32 keep 5.8314...), each band attenuated by
// exp(-rho l(l+1)/2) = exp(-rho {1,3,10,36}).
__device__ __forceinline__ void ide_features(const float d[3], float rho, float (&t)[48]) {
  const float x = d[0], y = d[1], z = d[2];
  const float x2 = x * x, y2 = y * y, z2 = z * z;
  const float xy = x * y, xz = x * z, yz = y * z;
  const float dxy = x2 - y2;
  const float a = 3.f * x2 - y2, b = x2 - 3.f * y2;
  const float z4 = z2 * z2, x4 = x2 * x2, y4 = y2 * y2;
  const float p = y4 - 10.f * x2 * y2 + 5.f * x4;
  const float q = x4 - 10.f * x2 * y2 + 5.f * y4;
  const float r = (x2 - 5.f * y2) * 7.f * x4 + (21.f * x2 - y2) * y4;
  const float s = (x2 - 21.f * y2) * x4 + (5.f * x2 - y2) * 7.f * y4;
  const float e1 = __expf(-rho), e2 = __expf(-3.f * rho), e4 = __expf(-10.f * rho), e8 = __expf(-36.f * rho);
  const float c1 = 0.48860251190291992f;
  t[0] = e1 * c1 * y;
  t[1] = e1 * c1 * z;
  t[2] = e1 * c1 * x;
  t[3] = e2 * 1.09254843059207907f * xy;
  t[4] = e2 * 1.09254843059207907f * yz;
  t[5] = e2 * 0.31539156525252001f * (3.f * z2 - 1.f);
  t[6] = e2 * 1.09254843059207907f * xz;
  t[7] = e2 * 0.54627421529603953f * dxy;
  t[8] = e4 * 2.50334294179670453f * xy * dxy;
  t[9] = e4 * 1.77013076977993053f * yz * a;
  t[10] = e4 * 0.94617469575756001f * xy * (7.f * z2 - 1.f);
  t[11] = e4 * 0.66904654355728916f * yz * (7.f * z2 - 3.f);
  t[12] = e4 * 0.1057855469152043038f * (35.f * z4 - 30.f * z2 + 3.f);
  t[13] = e4 * 0.66904654355728916f * xz * (7.f * z2 - 3.f);
  t[14] = e4 * 0.473087347878780009f * dxy * (7.f * z2 - 1.f);
  t[15] = e4 * 1.77013076977993053f * xz * b;
  t[16] = e4 * 0.62583573544917613f * (x2 * b - y2 * a);
  t[17] = e8 * 5.83141328139863895f * xy * (x2 * x4 - 7.f * x4 * y2 + 7.f * x2 * y4 - y2 * y4);
  t[18] = e8 * 5.83141328139863895f * yz * r;
  t[19] = e8 * 1.06466553211908514f * xy * (15.f * z2 - 1.f) * (3.f * x4 - 10.f * x2 * y2 + 3.f * y4);
  t[20] = e8 * 3.44991062209810801f * yz * (5.f * z2 - 1.f) * p;
  t[21] = e8 * 1.91366609903732278f * xy * (65.f * z4 - 26.f * z2 + 1.f) * dxy;
  t[22] = e8 * 1.23526615529554407f * yz * (39.f * z4 - 26.f * z2 + 3.f) * a;
  t[23] = e8 * 0.91230451686981894f * xy * (143.f * z4 * z2 - 143.f * z4 + 33.f * z2 - 1.f);
  t[24] = e8 * 0.1090412458987799555f * yz * (715.f * z4 * z2 - 1001.f * z4 + 385.f * z2 - 35.f);
  t[25] = e8 * 0.0090867704915649962938f * (6435.f * z4 * z4 - 12012.f * z4 * z2 + 6930.f * z4 - 1260.f * z2 + 35.f);
  t[26] = e8 * 0.1090412458987799555f * xz * (715.f * z4 * z2 - 1001.f * z4 + 385.f * z2 - 35.f);
  t[27] = e8 * 0.456152258434909470f * (143.f * z4 * z2 - 143.f * z4 + 33.f * z2 - 1.f) * dxy;
  t[28] = e8 * 1.23526615529554407f * xz * (39.f * z4 - 26.f * z2 + 3.f) * b;
  t[29] = e8 * 0.478416524759330697f * (65.f * z4 - 26.f * z2 + 1.f) * (x2 * b - y2 * a);
  t[30] = e8 * 3.44991062209810801f * xz * (5.f * z2 - 1.f) * q;
  t[31] = e8 * 0.53233276605954257f * (15.f * z2 - 1.f) * (x2 * q - y2 * p);
  t[32] = e8 * 5.83141328139863895f * xz * s;
  t[33] = e8 * 0.72892666017482986f * (x2 * s - y2 * r);
#pragma unroll
  for (int i = 34; i < 48; ++i) t[i] = 0.f;
}

}  // namespace rsnenc
