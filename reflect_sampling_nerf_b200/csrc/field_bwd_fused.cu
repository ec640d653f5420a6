// K5 backward in ONE launch: the dgrad chain (csrc/field_bwd_body.cuh) and the wgrad (csrc/field_wgrad_body.cuh) run
// side by side on disjoint CTAs of a single grid.  The chain is tensor / shared-memory bound, the wgrad HBM bound;
// run back to back (rsn_field_backward then rsn_field_wgrad) each leaves the other's resource idle, and every dY
// block makes a round trip through HBM.  Here a chain CTA publishes a per-tile flag once the tile's dY blocks have
// left shared memory, and the wgrad CTAs (which walk the tiles in the same increasing order) pick them up shortly
// after -- from L2 (126 MB, ~190 tiles of dY) rather than HBM.  The chain never waits on the wgrad and the grid is one
// wave of <= #SM CTAs (1 CTA per SM by shared memory), so the flag spin cannot deadlock.
// (validated, slower than the two-launch form -- 8.2 vs 6.4 ms at C2, DESIGN.md §4 -- and therefore part of the test
// build only: -DRSN_DEBUG_SWITCHES)
#ifdef RSN_DEBUG_SWITCHES
#include "field_bwd_body.cuh"
#include "field_wgrad_body.cuh"

namespace {

constexpr int F_THREADS = B_THREADS > W_THREADS ? B_THREADS : W_THREADS;

__global__ void __launch_bounds__(F_THREADS, 1)
    field_bwd_fused_kernel(const __grid_constant__ BwdParams bp, const __grid_constant__ WParams wp, const int n_chain,
                           int* tile_done) {
  if ((int)blockIdx.x < n_chain) {
    chain_body<KIND_BACKWARD>(bp, (int)blockIdx.x, n_chain, tile_done);
  } else {
    wgrad_body(wp, (int)blockIdx.x - n_chain, tile_done);
  }
}

}  // namespace

extern "C" int64_t rsn_field_backward_fused_workspace_bytes(int64_t n_points) {
  return ((n_points + TILE - 1) / TILE) * (int64_t)sizeof(int);
}

extern "C" int rsn_field_backward_fused(const void* wblob_t, const void* x_stash, int mode, const float* origins,
                                        const float* dirs, const float* area, const float* bins, int64_t n_rays,
                                        int64_t n_samples, const float* g_sigma, const float* g_feat, const float* feat,
                                        const float* aux, void* dy_stash, float* g_area, float* grad_blob,
                                        void* workspace, cudaStream_t stream) {
  RSN_ARG(mode == 0 || mode == 1, "rsn_field_backward_fused: mode must be 0 or 1");
  RSN_ARG(n_rays >= 0 && n_samples >= 1, "rsn_field_backward_fused: bad shape");
  if (n_rays == 0) return 0;
  RSN_ARG(n_rays * n_samples < (int64_t)2147483647 - TILE, "rsn_field_backward_fused: more than 2^31 points in one call");
  RSN_ARG(wblob_t && x_stash && dirs && g_feat && feat && aux && dy_stash && grad_blob && workspace,
          "rsn_field_backward_fused: null pointer");
  RSN_ARG(!g_area || mode == 1 || (origins && bins && area), "rsn_field_backward_fused: rays required for d pixel_area");
  RSN_ARG(((uintptr_t)wblob_t & 15) == 0 && ((uintptr_t)x_stash & 15) == 0 && ((uintptr_t)dy_stash & 15) == 0 &&
              ((uintptr_t)g_feat & 15) == 0 && ((uintptr_t)feat & 15) == 0 && ((uintptr_t)aux & 15) == 0,
          "rsn_field_backward_fused: buffers must be 16-byte aligned");
  BwdParams p = {};
  p.wblob_t = (const uint8_t*)wblob_t;
  p.x_stash = (const uint8_t*)x_stash;
  p.kind = KIND_BACKWARD;
  p.debug = rsn_env_int("RSN_BWD_DEBUG", 0);
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
  // CTA split: chain CTAs get the SM time the chain needs relative to the wgrad's MMA + load time (measured on the
  // separate kernels at C2: chain 3.0 ms x 148, wgrad 5.5 ms HBM-bound); RSN_FUSED_CHAIN_CTAS overrides.
  const int sms = rsn_num_sms();
  int n_chain = rsn_env_int("RSN_FUSED_CHAIN_CTAS", (sms * 54) / 100);
  n_chain = std::max(1, std::min(std::min(n_chain, sms - 14), p.n_tiles));
  WParams w;
  const int n_w = fill_wgrad_params(w, x_stash, dy_stash, p.n_points, grad_blob, sms - n_chain);
  RSN_CUDA(cudaMemsetAsync(workspace, 0, (size_t)p.n_tiles * sizeof(int), stream));
  const size_t smem = (size_t)std::max<size_t>(SM_TOTAL, (size_t)RING_BYTES) + 1024;
  static std::atomic<unsigned long long> done;
  RSN_CUDA(rsn_ensure_smem(field_bwd_fused_kernel, (int)smem, done));
  field_bwd_fused_kernel<<<n_chain + n_w, F_THREADS, smem, stream>>>(p, w, n_chain, (int*)workspace);
  RSN_LAUNCH_CHECK("field_bwd_fused_kernel");
  return 0;
}
#endif  // RSN_DEBUG_SWITCHES
