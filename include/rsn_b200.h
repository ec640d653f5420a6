/* rsn_b200 -- C-ABI of the B200-native (sm_100a) per-ray rendering hot path of
 * 236088/reflect-sampling-nerf.
 *
 * The reference is pure Python on un-vendored nerfstudio and has NO native interface of its own
 * (SURVEY.md 2.3); each entry point below names the reference call site(s) whose eager-PyTorch work it
 * replaces (paths relative to /root/reference/reflect_sampling_nerf/).  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference adds.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer to caller-owned memory (torch-allocated in the shipped host
 *     code); the library never allocates, frees or synchronises; work is enqueued on `stream`;
 *   - shapes are int64_t, tensors are dense row-major unless a stride argument says otherwise;
 *   - return value: 0 = ok, < 0 = bad argument, > 0 = cudaError_t; rsn_last_error() returns the
 *     thread-local message of the last failing call;
 *   - all entry points are re-entrant and CUDA-graph capturable; built with -arch=sm_100a only.
 */
#ifndef RSN_B200_H
#define RSN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* rsn_stream_t; /* == cudaStream_t */

/* ---- housekeeping ------------------------------------------------------------------------------ */
int rsn_version(void);               /* 100 = 0.1.0 */
const char* rsn_last_error(void);
int rsn_device_ok(void);             /* 0 iff the current device is sm_100 */

/* ---- K1: spaced ray sampling ---------------------------------------------------------------------
 * Replaces UniformSampler / ReciprocalSampler.generate_ray_samples:
 *   reflect_sampling_nerf_model.py:109,111,148,292 ; reflect_sampling_nerf_components.py:14-36.
 * lin_bins [S+1] = torch.linspace(0,1,S+1); t_rand [N, t_rand_cols] stratification noise or NULL (eval);
 * spacing_kind 0 = uniform, 1 = reciprocal (tan = 0.25).  Writes spacing and Euclidean bins [N,S+1].
 * Bit-exact with the oracle. */
int rsn_sample_spaced(const float* nears, const float* fars, const float* lin_bins, const float* t_rand,
                      int64_t t_rand_cols, int spacing_kind, float* spacing_bins, float* euclid_bins,
                      int64_t n_rays, int64_t n_samples, rsn_stream_t stream);

/* ---- K2: PDF (importance) resampling -------------------------------------------------------------
 * Replaces PDFSampler(include_original=False).generate_ray_samples:
 *   reflect_sampling_nerf_model.py:110,112,182,317.
 * weights [N,S] (row stride given), spacing_bins_in [N,S+1], u_base [S'+1] (eval: already centred),
 * rand [N,S'+1] or NULL.  Writes new spacing / Euclidean bins [N,S'+1] and, if inds_out != NULL, the
 * searchsorted(side="right") indices (int64) for the bit-exact index test. */
int rsn_pdf_resample(const float* weights, int64_t weights_row_stride, const float* spacing_bins_in,
                     const float* nears, const float* fars, const float* u_base, const float* rand,
                     int spacing_kind, float histogram_padding, float* spacing_bins_out,
                     float* euclid_bins_out, int64_t* inds_out, int64_t n_rays, int64_t n_in_samples,
                     int64_t n_out_samples, rsn_stream_t stream);

/* ---- K8: alpha compositing ------------------------------------------------------------------------
 * Replaces RaySamples.get_weights + Accumulation / RGB / Depth(median) / Normals / Semantic renderers:
 *   reflect_sampling_nerf_model.py:154-156,176,188-190,210,215-226,296,311,322,337,341.
 * sigma [N,S]; starts/ends point at element [0] of the per-ray start / end arrays with a common row
 * stride (pass bins and bins+1 with stride S+1 to composite straight from a bins array);
 * feat [N,S,C] per-sample channels, C in {0,1,3,4,8,16}.  Outputs: weights [N,S], accumulation [N],
 * depth_median [N], feat_out [N,C] = sum_s w*feat (background blending is the caller's per-ray op). */
int rsn_composite_fwd(const float* sigma, const float* starts, const float* ends, int64_t bin_row_stride,
                      const float* feat, int64_t n_channels, float* weights, float* accumulation,
                      float* depth_median, float* feat_out, int64_t n_rays, int64_t n_samples,
                      rsn_stream_t stream);
/* Backward of the above w.r.t. sigma and feat.  grad_weights [N,S], grad_accumulation [N],
 * grad_feat_out [N,C] may each be NULL (= zero).  grad_feat may be NULL (not needed). */
int rsn_composite_bwd(const float* sigma, const float* starts, const float* ends, int64_t bin_row_stride,
                      const float* feat, int64_t n_channels, const float* grad_weights,
                      const float* grad_accumulation, const float* grad_feat_out, float* grad_sigma,
                      float* grad_feat, int64_t n_rays, int64_t n_samples, rsn_stream_t stream);

/* 16-channel form (the model's feature row) with the two per-sample normal losses of
 * reflect_sampling_nerf_model.py:403-407 fused in: per ray pred_normal_loss = sum_s w_s |normals_s - feat_s[9:12]|^2 and
 * orientation_loss = sum_s w_s max(0, feat_s[13])^2 with w the (detached) compositing weights; normals [N,S,3].
 * The backward adds their gradients into grad_feat columns 9-11 and 13 (grad_* per-ray inputs may be NULL = 0). */
int rsn_composite16_fwd(const float* sigma, const float* starts, const float* ends, int64_t bin_row_stride,
                        const float* feat, const float* normals, float* weights, float* accumulation,
                        float* depth_median, float* feat_out, float* pred_normal_loss, float* orientation_loss,
                        int64_t n_rays, int64_t n_samples, rsn_stream_t stream);
int rsn_composite16_bwd(const float* sigma, const float* starts, const float* ends, int64_t bin_row_stride,
                        const float* feat, const float* normals, const float* grad_weights,
                        const float* grad_accumulation, const float* grad_feat_out,
                        const float* grad_pred_normal_loss, const float* grad_orientation_loss, float* grad_sigma,
                        float* grad_feat, int64_t n_rays, int64_t n_samples, rsn_stream_t stream);

/* ---- K3+K4+K5+K7: fused field forward --------------------------------------------------------------
 * Replaces, for one pass over the samples of a ray batch (mode 0), field.get_blob -> contract ->
 * get_density -> get_pred_normals / get_roughness / get_diff / get_tint -> IntegratedSHEncoding -> get_mid:
 *   reflect_sampling_nerf_field.py:90-186,203-207 ; reflect_sampling_nerf_components.py:52-140 ;
 *   call sites reflect_sampling_nerf_model.py:151-175,185-209,293-310,319-336
 * and, in mode 1, field.get_inf_color (reflect_sampling_nerf_field.py:190-201; model.py:290) with
 * dirs = w_r [N,3], area = sqradius [N], n_samples = 1 (origins/bins ignored).
 * wblob: bf16 weight blob of rsn_field_blob_bytes() bytes and bias: fp32 [rsn_field_bias_count()], both
 * produced by rsn_pack_field (or reflect_sampling_nerf_b200/packing.py).  origins/dirs [N,3], area [N]
 * (pixel_area), bins [N,S+1] Euclidean bin edges.  Outputs: sigma [N*S] (softplus density) and
 * feat [N*S,16]: 0-2 rgb = diff + tint*mid (mode 1: mid) | 3-5 diff | 6-8 tint | 9-11 pred_normal |
 * 12 sigmoid(roughness) | 13 n.d | 14 raw density | 15 softplus(roughness).
 * bf16 tensor-core MLP (tcgen05), fp32 encodings/heads. */
int rsn_field_forward(const void* wblob, const float* bias, int mode, const float* origins, const float* dirs,
                      const float* area, const float* bins, int64_t n_rays, int64_t n_samples, float* sigma,
                      float* feat, rsn_stream_t stream);
/* Training form of rsn_field_forward: additionally writes the activation stash (rsn_field_stash_bytes(N*S)
 * bytes: per 128-point tile 41 bf16 block images = IPE, the 8 hidden activations, bottleneck, IDE, mid hidden)
 * that the normals / backward / wgrad kernels read, and aux [N*S,8] = mid rgb (3), raw normal head (3), 2 spare. */
int rsn_field_forward_train(const void* wblob, const float* bias, int mode, const float* origins,
                            const float* dirs, const float* area, const float* bins, int64_t n_rays,
                            int64_t n_samples, float* sigma, float* feat, void* stash, float* aux,
                            rsn_stream_t stream);
int64_t rsn_field_stash_bytes(int64_t n_points);

/* ---- K6: density-gradient normals -------------------------------------------------------------------
 * Replaces Field.get_normals = -normalize(d raw_density / d contracted mean) (autograd.grad through the
 * density head, the 8 base layers and the IPE with the covariance held constant):
 *   reflect_sampling_nerf_field.py:125-127,134-135,146-147 ; reflect_sampling_nerf_model.py:159-160,194-195.
 * wblob_t: transposed bf16 weight blob (rsn_field_blob_t_bytes()), wd_bf16: the density head row as 256 bf16,
 * x_stash: the stash of the forward pass.  Output normals [N*S,3]. */
int rsn_field_normals(const void* wblob_t, const void* wd_bf16, const void* x_stash, int64_t n_rays,
                      int64_t n_samples, float* normals, rsn_stream_t stream);
int64_t rsn_field_blob_t_bytes(void);

/* ---- K5 backward, dgrad chain -----------------------------------------------------------------------
 * Replaces the autograd backward of reflect_sampling_nerf_field.py:122-186 (mode 0) / 190-201 (mode 1) for one
 * pass: from g_sigma [N*S] = dL/d sigma and g_feat [N*S,16] = dL/d feat (forward layout; columns 14,15 are
 * ignored; g_sigma may be NULL = zero) to the pre-activation gradient of every Linear, written to dy_stash (rsn_field_dy_stash_bytes) for
 * rsn_field_wgrad.  feat / aux are the forward outputs.  If g_area != NULL the chain continues through layer 0
 * and the IPE damping and writes dL/d pixel_area (mode 0) or dL/d sqradius (mode 1) of every POINT [N*S]
 * (the caller sums over the samples of a ray): the roughness -> cone width path of
 * reflect_sampling_nerf_model.py:272,286,290. */
int rsn_field_backward(const void* wblob_t, const void* x_stash, int mode, const float* origins, const float* dirs,
                       const float* area, const float* bins, int64_t n_rays, int64_t n_samples,
                       const float* g_sigma, const float* g_feat, const float* feat, const float* aux,
                       void* dy_stash, float* g_area, rsn_stream_t stream);
int64_t rsn_field_dy_stash_bytes(int64_t n_points);

/* ---- K5 backward, wgrad -----------------------------------------------------------------------------
 * dW = dY^T X and db = sum dY of every Linear of the field over all points of a pass, ACCUMULATED (fp32
 * atomics) into grad_blob, whose regions rsn_field_wgrad_layout describes (HOST pointers: offsets[2j] = dW
 * offset, offsets[2j+1] = db offset or -1, shapes[2j], shapes[2j+1] = rows, cols; returns the job count). */
int rsn_field_wgrad(const void* x_stash, const void* dy_stash, int64_t n_points, float* grad_blob,
                    rsn_stream_t stream);
int rsn_field_wgrad_layout(int64_t* host_offsets, int64_t* host_shapes, int64_t* total_floats);

/* Completes the gradient blob once per step, after the last rsn_field_wgrad (and after the all-reduce): the bottleneck
 * layer is linear in h7 and feeds only the mid layer, so the kernels stash neither its activations nor their gradients;
 * with G = dY_mid^T h7 (rows 64-191 of job 10's region) this fills dW/db of field_output_bottleneck (job 9's region) =
 * Wmb^T G, Wmb^T db_mid and job 12's region = the bottleneck columns of d mlp_mid.layers.0.weight = G Wb^T + db_mid bb^T
 * (and db_mid).
 * w_bott [256][256], b_bott [256], w_mid [128][290]: the fp32 parameters (field.py:54-86), device pointers. */
int rsn_field_wgrad_finish(float* grad_blob, const float* w_bott, const float* b_bott, const float* w_mid,
                           rsn_stream_t stream);
/* rsn_field_backward + rsn_field_wgrad in ONE launch: the dgrad-chain CTAs and the wgrad CTAs share the grid, the
 * chain publishes each tile's dY blocks with a per-tile flag (workspace: rsn_field_backward_fused_workspace_bytes,
 * zeroed by the call) and the wgrad picks them up from L2.  Same arguments and results as the two calls. */
int rsn_field_backward_fused(const void* wblob_t, const void* x_stash, int mode, const float* origins,
                             const float* dirs, const float* area, const float* bins, int64_t n_rays,
                             int64_t n_samples, const float* g_sigma, const float* g_feat, const float* feat,
                             const float* aux, void* dy_stash, float* g_area, float* grad_blob, void* workspace,
                             rsn_stream_t stream);
int64_t rsn_field_backward_fused_workspace_bytes(int64_t n_points);
/* Gradient blob -> flat fp32 gradient vector laid out parameter after parameter in rsn_pack_field order
 * (rsn_field_flat_layout: HOST array of 32 float offsets, returns the vector length, 617,742).  One launch. */
int rsn_unpack_grads(const float* grad_blob, float* flat_grads, rsn_stream_t stream);
int64_t rsn_field_flat_layout(int64_t* host_offsets32);

int64_t rsn_field_blob_bytes(void);
/* Packs the 32 fp32 parameter tensors of the field (HOST array of DEVICE pointers, order documented in
 * csrc/pack.cu: base weights 0-7, base biases 8-15, bottleneck, mid, rgb, density, normals, roughness, diff,
 * tint -- weight then bias each; names of reflect_sampling_nerf_field.py:54-86) into the forward blob, the
 * transposed blob, the bias vector and the bf16 density row.  Derived state: call after every optimizer step. */
int rsn_pack_field(const float* const* params, void* wblob, void* wblob_t, float* bias, void* wd_bf16,
                   rsn_stream_t stream);
int64_t rsn_field_bias_count(void);
/* The 16 IPE frequencies 2**linspace(0,16,16) the kernels use (HOST pointer; for the table test). */
int rsn_ipe_freqs(float* host_out16);

/* ---- K9: reflection set-up and composition of the bounce --------------------------------------------------
 * rsn_reflect_setup replaces reflect_sampling_nerf_model.py:215-229,267-271: from the composited feature row of the fine
 * pass comp16 [N,16] (rsn_field_forward feature layout), accumulation [N], median depth [N] and the rays, per ray:
 * diff [N,3] (white-blended), tint [N,3], normal [N,3] (safe-normalised), n_dot_d [N], mask [N] (uint8:
 * accumulation > 1e-2 and n.d < 0), bounce origins o + depth d and directions normalize(d - 2 (n.d) n).  clamp01 =
 * eval mode (RGBRenderer clamps).  The compaction of the masked rays stays with the caller (index list).
 * rsn_reflect_compose_* replace model.py:311-313,337-339: out = base; out[idx] = clip(diff[idx] + tint[idx] *
 * (comp[m,:3] + bg[m] (1 - acc[m])), 0, 1) and its backward w.r.t. comp, bg and base. */
int rsn_reflect_setup(const float* comp16, const float* acc, const float* depth, const float* origins, const float* dirs,
                      int clamp01, float* diff, float* tint, float* normal, float* n_dot_d, uint8_t* mask,
                      float* bounce_origins, float* bounce_dirs, int64_t n_rays, rsn_stream_t stream);
int rsn_reflect_compose_fwd(const float* base, const float* diff, const float* tint, const int64_t* idx,
                            const float* comp, int64_t comp_ld, const float* bg, const float* acc, int clamp_inner,
                            float* out, int64_t n_rays, int64_t n_bounced, rsn_stream_t stream);
int rsn_reflect_compose_bwd(const float* grad_out, const float* diff, const float* tint, const int64_t* idx,
                            const float* comp, int64_t comp_ld, const float* bg, const float* acc, float* grad_comp,
                            float* grad_bg, float* grad_base, int64_t n_rays, int64_t n_bounced, rsn_stream_t stream);

/* ---- tcgen05 building-block probes (unit tests of csrc/umma.cuh) ---------------------------------- */
int rsn_probe_umma_kmajor(const void* x_blocks, const void* w_blocks, int64_t n_out, int64_t k_blocks,
                          int64_t n_split, float* out, rsn_stream_t stream);
int rsn_probe_umma_mnmajor(const void* u_blocks, const void* v_blocks, int64_t m_blocks, int64_t n_blocks,
                           float* out, rsn_stream_t stream);

/* A-from-TMEM probe (tcgen05.mma [d], [a], b-desc): out [128, n_out] = X * W^T with X staged into TMEM by tcgen05.st.
 * iters > 0 also times `iters` back-to-back MMAs of that form into *cycles_out (DEVICE int64). */
int rsn_probe_umma_ts(const void* x_blocks, const void* w_blocks, int64_t n_out, int64_t k_blocks, float* out,
                      int64_t iters, int64_t* cycles_out, rsn_stream_t stream);

/* TMEM read / write throughput seen by n_warps (1..8) epilogue-style warps, 64 columns x 32 lanes per iteration and warp,
 * optionally while another warp streams mma_iters tcgen05.mma into the other accumulator buffer; cycles_out: DEVICE
 * int64 [9], cycles of each reader warp for `iters` iterations and ([8]) of the MMA stream.  mode: see csrc/probe.cu. */
int rsn_probe_tmem_rate(int64_t n_warps, int64_t mode, int64_t iters, int64_t mma_iters, int64_t* cycles_out,
                        rsn_stream_t stream);

/* Step-by-step cost of the field kernels' epilogue (four warps, one 64-column group per iteration); `steps` is a bit set
 * of the stages to include (csrc/probe.cu); cycles_out as for rsn_probe_tmem_rate. */
int rsn_probe_epilogue(int64_t steps, int64_t iters, int64_t mma_iters, int64_t* cycles_out, rsn_stream_t stream);

/* CTA-pair probe: out [256, n_out] = X [256, 64 k_blocks] * W [n_out, 64 k_blocks]^T with tcgen05.mma.cta_group::2;
 * x_blocks = two tiles of k_blocks block images, w_blocks = k_blocks images of n_out rows. */
int rsn_probe_umma_2cta(const void* x_blocks, const void* w_blocks, int64_t n_out, int64_t k_blocks, float* out,
                        rsn_stream_t stream);
/* Issue-rate probe: cycles for `iters` back-to-back M128 x n x K16 bf16 tcgen05.mma with K-major (0) or
 * MN-major (1) A / B operands; *cycles_out is a DEVICE int64. */
int rsn_probe_umma_rate(int a_major, int b_major, int64_t n, int64_t iters, int64_t* cycles_out, rsn_stream_t stream);

/* Same for the CTA pair: cycles for `iters` back-to-back M256 x n x K16 cta_group::2 MMAs on n_pairs clusters. */
int rsn_probe_umma_rate_2cta(int64_t n, int64_t iters, int64_t n_pairs, int64_t* cycles_out, rsn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RSN_B200_H */
