// K5 backward (dgrad chain) and K6 (density-gradient normals) of the fused field, for 128-point tiles.
// SURVEY.md §2.4 (K5 bwd, K6, K3/K4 bwd for the reflected passes), §8 rows a8-a10, a12, a17; App. D.
//
// Replaces, per pass over the samples of a ray batch:
//   * chain kind NORMALS: Field.get_normals = -normalize(d raw_density / d contracted mean), cov held constant
//       reflect_sampling_nerf_field.py:125-127,134-135,146-147 ; model.py:159-160,194-195
//   * chain kind BACKWARD: the autograd backward of field.py:122-186 (+190-201 for the infinity colour) from the
//       per-sample gradients (d sigma, d rgb, d pred_normal, d n.d, d sigmoid roughness) down to the
//       pre-activation gradient dY_l of every Linear, which it leaves in HBM as bf16 block images for the
//       wgrad kernel (csrc/field_wgrad.cu); for the reflected passes and the infinity colour it continues
//       through layer 0 and the IPE damping to d pixel_area / d sqradius (the roughness -> cone-width path,
//       model.py:272,286,290).
//
// Same machinery as the forward (csrc/field_fwd.cu): persistent CTA per SM, one tile in flight; warp 0
// streams the TRANSPOSED weight blob (B[n = input feature][k = output feature]) through a 5 x 32 KB ring,
// warp 6 streams the two stashed encoding blocks the IPE Jacobian needs through a 2 x 16 KB ring, warp 1 issues
// tcgen05.mma into two alternating 256-column TMEM accumulators, warps 2-5 run the per-row epilogues:
// dX = acc * (h > 0) -> bf16 -> the next step's A operand, published per 64-column group and written back into TMEM
// (tcgen05.st; the next MMA takes its A operand from there).  The ReLU masks arrive as the forward's bit masks (8 bytes
// per row, layer and group, prefetched into registers).  BACKWARD launches four more warps (7-10) that read every
// handed-over dY group back out of TMEM and write it to the dY stash with coalesced st.global (chunk-major block image,
// csrc/field_layout.cuh): the stash is not on the step-critical path, and the issuer re-uses an accumulator buffer only
// after they have released it.  The test build keeps the shared-memory operand form (RSN_BWD_TS=0; its epilogue writes
// the stash rows itself), pinned bit-identical by tests.
//
// Roofline: bf16 tensor.  Algorithmic FLOPs per point: NORMALS 1,019,392; BACKWARD dgrad 1,179,904 primary,
// 1,229,056 reflected.  HBM: masks 288 B/pt (+ encodings 256 B/pt) in, dY 4.4 KB/pt out (dY of the
// bottleneck layer is not written: csrc/field_wgrad.cu).
#include "field_bwd_body.cuh"


extern "C" int64_t rsn_field_blob_t_bytes(void) { return (int64_t)BWD_BLOB_BYTES; }

extern "C" int rsn_field_normals(const void* wblob_t, const void* wd_bf16, const void* x_stash, int64_t n_rays,
                                 int64_t n_samples, float* normals, cudaStream_t stream) {
  RSN_ARG(n_rays >= 0 && n_samples >= 1, "rsn_field_normals: bad shape");
  if (n_rays == 0) return 0;
  RSN_ARG(n_rays * n_samples < (int64_t)2147483647 - TILE, "rsn_field_normals: more than 2^31 points in one call");
  RSN_ARG(wblob_t && wd_bf16 && x_stash && normals, "rsn_field_normals: null pointer");
  RSN_ARG(((uintptr_t)wblob_t & 15) == 0 && ((uintptr_t)wd_bf16 & 15) == 0 && ((uintptr_t)x_stash & 15) == 0,
          "rsn_field_normals: blobs must be 16-byte aligned");
  BwdParams p = {};
  p.wblob_t = (const uint8_t*)wblob_t;
  p.x_stash = (const uint8_t*)x_stash;
  p.kind = KIND_NORMALS;
  p.n_samples = (int)n_samples;
  p.n_points = (int)(n_rays * n_samples);
  p.n_tiles = (p.n_points + TILE - 1) / TILE;
  p.wd_bf16 = (const uint32_t*)wd_bf16;
  p.normals = normals;
  return launch_chain<KIND_NORMALS>(p, stream);
}

extern "C" int rsn_field_backward(const void* wblob_t, const void* x_stash, int mode, const float* origins,
                                  const float* dirs, const float* area, const float* bins, int64_t n_rays,
                                  int64_t n_samples, const float* g_sigma, const float* g_feat, const float* feat,
                                  const float* aux, void* dy_stash, float* g_area, const int* n_rays_dev,
                                  cudaStream_t stream) {
  RSN_ARG(mode == 0 || mode == 1, "rsn_field_backward: mode must be 0 or 1");
  RSN_ARG(n_rays >= 0 && n_samples >= 1, "rsn_field_backward: bad shape");
  if (n_rays == 0) return 0;
  RSN_ARG(n_rays * n_samples < (int64_t)2147483647 - TILE, "rsn_field_backward: more than 2^31 points in one call");
  RSN_ARG(wblob_t && x_stash && dirs && g_feat && feat && aux && dy_stash, "rsn_field_backward: null pointer");
  RSN_ARG(!g_area || mode == 1 || (origins && bins && area), "rsn_field_backward: rays required for d pixel_area");
  RSN_ARG(((uintptr_t)wblob_t & 15) == 0 && ((uintptr_t)x_stash & 15) == 0 && ((uintptr_t)dy_stash & 15) == 0 &&
              ((uintptr_t)g_feat & 15) == 0 && ((uintptr_t)feat & 15) == 0 && ((uintptr_t)aux & 15) == 0,
          "rsn_field_backward: buffers must be 16-byte aligned");
  BwdParams p = {};
  p.wblob_t = (const uint8_t*)wblob_t;
  p.x_stash = (const uint8_t*)x_stash;
  p.kind = KIND_BACKWARD;
  p.debug = rsn_env_int("RSN_BWD_DEBUG", 0);
  p.n_rays_dev = n_rays_dev;
  p.mode = mode;
  p.want_area = g_area != nullptr;
  p.origins = origins;
  p.dirs = dirs;
  p.area = area;
  p.bins = bins;
  p.n_samples = (int)n_samples;
  p.n_points = (int)(n_rays * n_samples);
  p.n_tiles = (p.n_points + TILE - 1) / TILE;
  p.g_sigma = g_sigma;
  p.g_feat = g_feat;
  p.feat = feat;
  p.aux = aux;
  p.dy_stash = (uint8_t*)dy_stash;
  p.g_area = g_area;
  return launch_chain<KIND_BACKWARD>(p, stream);
}
