/* rsn_b200_test.h -- entry points that exist ONLY in the test build librsn_b200_dbg.so
 * (python -m reflect_sampling_nerf_b200.build compiles csrc/ a second time with -DRSN_DEBUG_SWITCHES + csrc/probe.cu).
 * That build also honours the RSN_FWD_TS (inference) / RSN_BWD_TS / RSN_*_DEBUG environment switches (alternative operand
 * forms that must stay bit-identical, timing ablations whose results are wrong by design); the product library
 * librsn_b200.so (include/rsn_b200.h) contains none of this. */
#ifndef RSN_B200_TEST_H
#define RSN_B200_TEST_H
#include "rsn_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* In-kernel cycle trace of the forward field kernel (RSN_FWD_DEBUG & 32): (tag, clock64) pairs of the third tile of CTA 0 of
 * the last traced launch -> HOST host_out[2 * max_pairs]; returns the number of pairs, resets the trace, synchronises.
 * Tags: csrc/field_fwd.cu. */
int rsn_debug_fwd_trace(long long* host_out, int max_pairs);

/* ---- tcgen05 building-block probes (unit tests of csrc/umma.cuh) ---------------------------------- */
int rsn_probe_umma_kmajor(const void* x_blocks, const void* w_blocks, int64_t n_out, int64_t k_blocks,
                          int64_t n_split, float* out, rsn_stream_t stream);
int rsn_probe_umma_mnmajor(const void* u_blocks, const void* v_blocks, int64_t m_blocks, int64_t n_blocks,
                           float* out, rsn_stream_t stream);
/* The wgrad's operand form for CHUNK-MAJOR stash blocks (csrc/field_layout.cuh): out [128, 64 n_blocks] = U^T V with U
 * [128 points, 128] and V [128 points, 64 n_blocks] given as chunk-major block images and read through NO-swizzle MN-major
 * descriptors whose LBO / SBO fields are passed in (128 / 1024 is the pair the wgrad uses).  iters > 0 also times the MMA
 * stream into *cycles_out (DEVICE int64). */
int rsn_probe_umma_mnmajor_cm(const void* u_blocks, const void* v_blocks, int64_t n_blocks, int64_t lbo, int64_t sbo,
                              float* out, int64_t iters, long long* cycles_out, rsn_stream_t stream);

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
#endif /* RSN_B200_TEST_H */
