// Weight / bias gradients of the field: dW_l = dY_l^T X_l and db_l = sum_points dY_l, accumulated over every
// 128-point tile of a pass.  SURVEY.md §2.4 K5 (bwd, wgrad), §8 rows a8, a9, a12.
//
// Replaces the autograd wgrad of every nn.Linear the reference evaluates per sample:
//   mlp_base.layers.0-7, field_output_{density,normals,roughness,diff,tint,bottleneck}, mlp_mid.layers.0,
//   field_output_mid            reflect_sampling_nerf_field.py:54-86 (forward at field.py:132-186)
//
// Operands are the block images the forward (X, csrc/field_fwd.cu stash) and the dgrad chain (dY,
// csrc/field_bwd.cu) left in HBM: [128 points][64 features] bf16, 128-byte swizzled.  Read "transposed"
// (MN-major UMMA descriptors, features contiguous) they are directly the A = dY^T and B = X operands of
//   D[out, in] (+)= sum over 16 points  dY[pt, out] * X[pt, in]            (tcgen05.mma, fp32 in TMEM)
// One CTA owns one job = (layer, all <=256 output features, <=256 input features) and a contiguous range of
// tiles (split-K over points); its accumulators stay in TMEM for the whole range and are flushed once with
// fp32 atomics.  db comes from the same shared-memory dY slabs, summed by the otherwise idle epilogue warps.
//
// HBM-bound: algorithmic bytes = (m_blocks + n_blocks) * 16 KB per tile and job; 95 blocks = 1.52 MB per tile
// over the 12 jobs (11.9 KB per point) against 2 * 618k * 128 = 158 MFLOP per tile.
#include "field_wgrad_body.cuh"


extern "C" int64_t rsn_field_dy_stash_bytes(int64_t n_points) {
  return ((n_points + TILE - 1) / TILE) * (int64_t)DY_BLOCKS * BLOCK_BYTES;
}

// Float offsets of every job's dW / db region inside the gradient blob: out[2*j] = dW offset, out[2*j+1] = db
// offset or -1.  Returns the number of jobs; *total_floats = blob size.  HOST pointers.
extern "C" int rsn_field_wgrad_layout(int64_t* host_offsets, int64_t* host_shapes, int64_t* total_floats) {
  int64_t off = 0;
  for (int j = 0; j < kNumJobs; ++j) {
    const int m = kJobs[j].m_blocks * 64, n = kJobs[j].n_blocks * 64;
    if (host_offsets) host_offsets[2 * j] = off;
    if (host_shapes) host_shapes[2 * j] = m, host_shapes[2 * j + 1] = n;
    off += (int64_t)m * n;
    if (host_offsets) host_offsets[2 * j + 1] = kJobs[j].has_db ? off : -1;
    if (kJobs[j].has_db) off += m;
  }
  if (total_floats) *total_floats = off;
  return kNumJobs;
}

extern "C" int rsn_field_wgrad(const void* x_stash, const void* dy_stash, int64_t n_points, float* grad_blob,
                               cudaStream_t stream) {
  RSN_ARG(n_points >= 0, "rsn_field_wgrad: bad shape");
  if (n_points == 0) return 0;
  RSN_ARG(x_stash && dy_stash && grad_blob, "rsn_field_wgrad: null pointer");
  RSN_ARG(((uintptr_t)x_stash & 15) == 0 && ((uintptr_t)dy_stash & 15) == 0, "rsn_field_wgrad: stashes must be 16-byte aligned");
  WParams p;
  const int cta = fill_wgrad_params(p, x_stash, dy_stash, n_points, grad_blob, rsn_num_sms());
  const size_t smem = (size_t)RING_BYTES + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    RSN_CUDA(cudaFuncSetAttribute(field_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  field_wgrad_kernel<<<cta, W_THREADS, smem, stream>>>(p);
  RSN_LAUNCH_CHECK("field_wgrad_kernel");
  return 0;
}
