// K3+K4+K5+K7 fused: frustum -> Gaussian -> contraction -> integrated positional encoding -> 8x256 base
// MLP -> heads -> integrated directional encoding -> bottleneck / mid MLP -> rgb, for 128-point tiles.
// SURVEY.md §2.4 (K3,K4,K5,K7), §8 rows a5-a9, a11, a12, a15, a17.
//
// Replaces (paths relative to /root/reference/reflect_sampling_nerf/):
//   field.get_blob / contract / get_density          reflect_sampling_nerf_field.py:90-137
//   get_pred_normals / get_roughness / get_diff / get_tint / get_mid / get_reflection (n.d)
//                                                     reflect_sampling_nerf_field.py:139-207
//   get_inf_color (mode 1)                            reflect_sampling_nerf_field.py:190-201
//   IntegratedSHEncoding                              reflect_sampling_nerf_components.py:52-140
//   call sites                                        reflect_sampling_nerf_model.py:151-175,185-209,290,293-310,319-336
//
// One persistent CTA per SM, 10 warps (14 in training), one 128-point tile in flight per CTA:
//   warp 0      weight producer: streams the pre-packed bf16 weight blob (L2 resident, 1.27 MB) through a ring of
//               5 x 32 KB shared-memory stages with cp.async.bulk + mbarrier complete_tx
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M128, N<=256, K16, bf16 -> fp32 in TMEM) and requests each
//               wide layer's fp32 bias into a two-slot shared-memory buffer two layers ahead
//   warps 6-9   epilogue: tcgen05.ld the accumulator row of "their" point, bias + ReLU, bf16, and hand the row to the next
//               layer as its A operand -- by default back into TMEM (tcgen05.st over accumulator columns already read;
//               the next MMA is tcgen05.mma [d], [a], b-desc), optionally (RSN_FWD_TS=0, inference only) in place into the
//               swizzled shared-memory activation blocks; heads, IDE, outputs
//   warps 2-5   prologue for the NEXT tile: frustum gaussian, contraction, IPE -> bf16 A operand (shared memory)
//   warps 10-13 (training launches only) stash: read the handed-over bf16 operand back out of TMEM (it stays valid until the
//               issuer re-uses the accumulator buffer two layers later, which waits for a_free), derive the ReLU bit masks
//               and write the row straight to the activation stash with coalesced st.global.v4 (chunk-major block image,
//               field_layout.cuh) -- no shared-memory staging, no proxy fence, nothing on the layer-critical path
//               (a layer's groups after its LAST hand-over, RSN_STASH_LAG_FWD; training layers are plain N = 256 steps)
// The two 256-column TMEM accumulator buffers alternate by layer, and every layer's epilogue publishes
// its output per 64-column group (mbarrier act_ready[g]) so that the next layer's K-block g is issued as
// soon as that group is written: MMA of layer l+1 overlaps the epilogue of layer l.
//
// Roofline: bf16 tensor.  Algorithmic FLOPs per point: 1,230,592 (primary),
// 1,229,056 (reflected), 1,225,472 (infinity colour).  HBM: 24 B/ray + 8 B/sample in, 68 B/point out (+ 5,024 B/point of
// stash in training).
#include "rsn_common.cuh"
#include "umma.cuh"
#include "field_layout.cuh"
#include "encodings.cuh"
#include <algorithm>
#include <type_traits>
#include <stdlib.h>
#include <atomic>

namespace {

using namespace umma;
using namespace rsnf;
using namespace rsnenc;

constexpr int NUM_STAGES = 3;
constexpr int SMEM_ENC = 0;                                    // 2 buffers x 2 blocks
constexpr int SMEM_ACT = 4 * BLOCK_BYTES;                      // 4 blocks (TS inference: two more ring stages)
constexpr int SMEM_W = SMEM_ACT + 4 * BLOCK_BYTES;             // ring
constexpr int SMEM_BARS = SMEM_W + NUM_STAGES * W_STAGE_BYTES;   // 229,376: mbarriers (256 B)
// two bias slots: [0, 1024) fp32 bias of a wide layer | [1024, 1088) the 16 head biases (arrive with the bottleneck layer's)
// | [1088, 1152) the rgb layer's (arrive with the mid layer's)
constexpr int BIAS_SLOT_BYTES = 1152;
constexpr int SMEM_BIAS = SMEM_BARS + 256;
constexpr int SMEM_TOTAL = SMEM_BIAS + 2 * BIAS_SLOT_BYTES;      // 231,936 of the 232,448 a CTA may have
constexpr int WIDE_LAYERS = 10;                                  // per tile: base 0..7, bottleneck, mid
constexpr int NUM_THREADS = 320;          // inference launches
#ifndef RSN_FWD_TRAIN_SPLIT
#define RSN_FWD_TRAIN_SPLIT 0
#endif
#ifndef RSN_STASH_LAG_FWD
#define RSN_STASH_LAG_FWD 4
#endif
constexpr int NUM_THREADS_TRAIN = 448;    // + the four stash warps

// (Biases never live in __constant__ memory: every layer's fp32 bias reaches the epilogue through a two-slot shared-memory
// buffer filled by bulk copies from the caller's bias vector, so two launches with different parameters on two streams
// share no state -- and an indexed LDC of a 10 KB table that cycles once per tile was the forward's bottleneck)
const float h_freq[16] = {
    0x1.0000000000000p+0f,  0x1.0c1b780000000p+1f,  0x1.18c9880000000p+2f,  0x1.26111c0000000p+3f,
    0x1.33f9760000000p+4f,  0x1.428a320000000p+5f,  0x1.51cb4e0000000p+6f,  0x1.61c5140000000p+7f,
    0x1.7280340000000p+8f,  0x1.8405f60000000p+9f,  0x1.965fde0000000p+10f, 0x1.a998080000000p+11f,
    0x1.bdb8d20000000p+12f, 0x1.d2cd4c0000000p+13f, 0x1.e8e1020000000p+14f, 0x1.0000000000000p+16f};

struct FwdParams {
  const uint8_t* wblob;     // forward weight blob (FWD_BLOB_BYTES)
  const float* bias;        // [N_BIAS]
  int mode;                 // 0 = frustum samples of a ray batch, 1 = infinity colour (one point per ray)
  const float* origins;     // [N,3]   (mode 0)
  const float* dirs;        // [N,3]
  const float* area;        // [N] pixel_area (mode 0) | sqradius (mode 1)
  const float* bins;        // [N,S+1] euclidean bins (mode 0)
  int n_samples;            // S (mode 1, 2: 1)
  int n_points;             // P = N*S (capacity when n_rays_dev is set)
  int n_tiles;
  const int* n_rays_dev;    // optional device-side ray count: the launch covers min(*n_rays_dev, N) rays
  // mode 2 (the Field method API, field.py:122-186): one point per "ray" with a caller-supplied Gaussian
  const float* pt_mean;     // [P,3]   (contracted) mean
  const float* pt_cov;      // [P,3,3] (contracted) covariance; only its diagonal is read (NeRFEncoding, App. A.4)
  const float* pt_rho;      // [P] roughness fed to the IDE instead of softplus(roughness head), or NULL
  float* sigma;             // [P]
  float* feat;              // [P][16]
  uint8_t* stash;           // training: [n_tiles][STASH_BLOCKS][16 KB] activation block images, or NULL
  float* aux;               // training: [P][8] = mid rgb (3), raw normal head (3), raw roughness head, 1 spare, or NULL
  int debug;                // RSN_FWD_DEBUG (test build only; timing experiments): 2 = no trig in the prologue, 4 = no weight streaming, 8 = no stash bulk stores (TS form)
};

// In-kernel cycle trace (test build, RSN_FWD_DEBUG & 32): (tag, clock64) pairs of the third tile of CTA 0, read back with
// rsn_debug_fwd_trace.  Tags: 1000 + 10 l + g issuer saw act_ready[g] while issuing layer l | 1500 + l issuer committed
// layer l | 2000 + l epilogue (warp 6, lane 0) saw acc_full of layer l | 2100 + 10 l + g group g handed over |
// 2200 + 10 l + g group g staged (stash) | 3000 prologue row encoded.
#ifdef RSN_DEBUG_SWITCHES
// fire-and-forget stores into a per-role region (issuer: pairs 0..2047, epilogue: 2048..4095) with the index in a register:
// nothing on the traced warp's critical path but the clock read and one st.global
__device__ long long g_fwd_trace[8192];
__device__ int g_fwd_trace_n[2];
#define RSN_TRACE(on, tag)                                            \
  do {                                                                \
    if ((on) && trace_i < 2047) {                                     \
      g_fwd_trace[trace_base + 2 * trace_i] = (long long)(tag);       \
      g_fwd_trace[trace_base + 2 * trace_i + 1] = clock64();          \
      ++trace_i;                                                      \
    }                                                                 \
  } while (0)
#define RSN_TRACE_DECL(role) int trace_i = 0; const int trace_base = (role) * 4096; (void)trace_base
#define RSN_TRACE_END(on, role) do { if (on) g_fwd_trace_n[role] = trace_i; } while (0)
#else
#define RSN_TRACE(on, tag) \
  do {                     \
  } while (0)
#define RSN_TRACE_DECL(role) \
  do {                       \
  } while (0)
#define RSN_TRACE_END(on, role) \
  do {                          \
  } while (0)
#endif

constexpr int MAX_STAGES = 5;
struct Barriers {
  uint64_t w_full[MAX_STAGES], w_empty[MAX_STAGES];
  uint64_t enc_full[2], enc_empty[2];
  uint64_t act_ready[8];         // [accumulator buffer the producing layer used][64-column group]: two sets, so that the
                                 // stash warps may lag the issuer by up to two layers without aliasing the phase parity
  uint64_t a_free[2];            // training: the stash warps have read the operand that lived in accumulator buffer b
  uint64_t ide_ready;
  uint64_t acc_full[4];          // [accumulator buffer][output half: columns 0-127 | 128-255]
  uint64_t bias_full[2];         // the bias slot (wide layer j -> slot j & 1) has landed
  uint32_t tmem_slot;
};
static_assert(sizeof(Barriers) <= 256, "Barriers must fit their shared-memory slot");

// ---------------------------------------------------------------------------------------------- helpers
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 8 encoded columns -> one 16-byte chunk of the row
__device__ __forceinline__ void store_chunk(uint32_t row_saddr, int row, int chunk, const float (&f)[8]) {
  sts128(row_saddr + (uint32_t)((chunk ^ (row & 7)) << 4), pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
         pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// ---------------------------------------------------------------------------------------------- prologue
// Contracted Gaussian of one frustum sample: field.get_blob (field.py:90-96) followed by
// field.contract (field.py:98-119).  Only diag(J cov J) is consumed downstream (NeRFEncoding reads
// torch.diagonal(covs) only).  The mean follows the reference's fp32 operation order without fma
// contraction because it feeds sin(2 pi x f) at f up to 65536.
__device__ __forceinline__ void frustum_gaussian_contracted(const float o[3], const float d[3], float t0, float t1,
                                                            float pixel_area, float (&xm)[3], float (&dg)[3]) {
  const float radius = __fdiv_rn(__fsqrt_rn(pixel_area), 1.7724538509055159f);
  const float mu = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
  const float hw = __fmul_rn(__fsub_rn(t1, t0), 0.5f);
  const float hw2 = __fmul_rn(hw, hw), mu2 = __fmul_rn(mu, mu);
  const float den = __fadd_rn(__fmul_rn(3.0f, mu2), hw2);
  const float tmean = __fadd_rn(mu, __fdiv_rn(__fmul_rn(__fmul_rn(2.0f, mu), hw2), den));
  float m[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) m[a] = __fadd_rn(o[a], __fmul_rn(d[a], tmean));
  const float hw4 = hw2 * hw2;
  const float dir_var = hw2 / 3.0f - 0.26666668f * ((hw4 * (12.0f * mu2 - hw2)) / (den * den));
  const float rad_var = (radius * radius) * (mu2 * 0.25f + 0.41666666f * hw2 - 0.26666668f * hw4 / den);
  // cov = dir_var d d^T + rad_var (I - d (d / max(|d|^2, 1e-10))^T)      (symmetric; 6 entries)
  const float dd = fmaxf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2], 1e-10f);
  const float dn[3] = {d[0] / dd, d[1] / dd, d[2] / dd};
  float cov[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      cov[i][j] = dir_var * (d[i] * d[j]) + rad_var * ((i == j ? 1.0f : 0.0f) - d[i] * dn[j]);
  // contraction
  const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(m[0], m[0]), __fmul_rn(m[1], m[1])), __fmul_rn(m[2], m[2]));
  const float n1 = __fsqrt_rn(n2);
  if (n1 > 1.0f) {
    const float sc = __fdiv_rn(__fsub_rn(__fmul_rn(2.0f, n1), 1.0f), n2);
    // J = ((2 n1 - 2)(I - m m^T / n2) + I) / n2   (symmetric)
    const float a2 = 2.0f * n1 - 2.0f;
    float J[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float e = (i == j) ? 1.0f : 0.0f;
        J[i][j] = (a2 * (e - m[i] * m[j] / n2) + e) / n2;
      }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      // (J cov J)_ii = sum_k (sum_j J_ij cov_jk) J_ki
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) t += J[i][j] * cov[j][k];
        acc += t * J[k][i];
      }
      dg[i] = fmaxf(acc, 0.f);
      xm[i] = __fmul_rn(sc, m[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      dg[i] = fmaxf(cov[i][i], 0.f);
      xm[i] = m[i];
    }
  }
}

// IPE of (xm, diag dg) -> bf16 row of the enc operand (columns 0..111; 99..111 zero).
// NeRFEncoding.forward with covs: s = fl(fl(2pi x) f), v = fl(diag fl(f f)),
// enc = exp(-v/2) sin(s | s + pi/2), raw xyz appended last.
__device__ __forceinline__ void encode_row(uint32_t enc_saddr, int row, const float (&xm)[3], const float (&dg)[3],
                                           bool zero_tail) {
  const uint32_t row0 = enc_saddr + (uint32_t)row * 128u;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float sx = __fmul_rn(6.2831854820251465f, xm[a]);
      const float va = dg[a];
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        float f8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          f8[i] = ipe_value(sx, va, c_freq[kk * 8 + i], half);
        }
        const int j = half * 6 + a * 2 + kk;  // 16-byte chunk index along the 112 columns
        store_chunk(row0 + (uint32_t)(j >> 3) * BLOCK_BYTES, row, j & 7, f8);
      }
    }
  }
  const float f12[8] = {xm[0], xm[1], xm[2], 0.f, 0.f, 0.f, 0.f, 0.f};
  const float f13[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  store_chunk(row0 + BLOCK_BYTES, row, 12 & 7, f12);
  store_chunk(row0 + BLOCK_BYTES, row, 13 & 7, f13);
  if (zero_tail) {  // columns 112..127 are never read by the forward MMAs but the wgrad reads whole blocks
    store_chunk(row0 + BLOCK_BYTES, row, 14 & 7, f13);
    store_chunk(row0 + BLOCK_BYTES, row, 15 & 7, f13);
  }
}

// ---------------------------------------------------------------------------------------------- epilogue
// 64 accumulator columns of this thread's row (+bias, optional ReLU) -> bf16 -> the next layer's A operand.
// The 64 biases come from shared memory (bias_saddr; every lane reads the same 16 bytes: one broadcast wavefront per
// load).  (A __constant__ table was an INDEXED load -- the layer is a run-time value -- of 10 KB that cycles once per
// tile: 32 LDC per group, 0.5 ms of 2.66 at C2, measured by replacing the bias with a register constant.)
// TS: the packed row goes back into TMEM (columns a_taddr..+31, over accumulator columns this thread has already read);
// otherwise (SS form, inference in the test build) it is written in place into the swizzled shared-memory block.
template <bool RELU, bool TS>
__device__ __forceinline__ void epilogue_group(uint32_t tmem_row_col, uint32_t blk_saddr, int row, uint32_t bias_saddr,
                                               uint32_t a_taddr) {
  uint32_t v[2][32];
  tmem_ld32(tmem_row_col, v[0]);
  tmem_ld32(tmem_row_col + 32, v[1]);
  float4 b[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(b[i].x), "=f"(b[i].y), "=f"(b[i].z), "=f"(b[i].w)
                 : "r"(bias_saddr + (uint32_t)i * 16u));
  }
  tmem_ld_wait();
  const uint32_t row_saddr = blk_saddr + (uint32_t)row * 128u;
  uint32_t a[32];   // TS: the 64 bf16 of this row, two per 32-bit TMEM column
#pragma unroll
  for (int h = 0; h < 2; ++h) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {  // 8 columns per 16-byte chunk
      uint32_t pk[4];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 bb = b[h * 8 + c * 2 + q];
        const float x0 = __uint_as_float(v[h][c * 8 + q * 4 + 0]) + bb.x;
        const float x1 = __uint_as_float(v[h][c * 8 + q * 4 + 1]) + bb.y;
        const float x2 = __uint_as_float(v[h][c * 8 + q * 4 + 2]) + bb.z;
        const float x3 = __uint_as_float(v[h][c * 8 + q * 4 + 3]) + bb.w;
        pk[q * 2 + 0] = RELU ? pack_relu_bf16x2(x0, x1) : pack_bf16x2(x0, x1);
        pk[q * 2 + 1] = RELU ? pack_relu_bf16x2(x2, x3) : pack_bf16x2(x2, x3);
      }
      const int chunk = h * 4 + c;
      if (!TS) sts128(row_saddr + (uint32_t)((chunk ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
      if (TS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) a[chunk * 4 + j] = pk[j];
      }
    }
  }
  if (TS) tmem_st32(a_taddr, a);
}

// Stash warps: the packed row (32 words = 64 bf16 of one 64-column group) -> its place in the stash block (chunk-major image:
// a warp-level st.global.v4 writes 512 contiguous bytes) and the ReLU bit masks of the group (one packed compare + one LOP3
// per word; word i of a 32-column half contributes bits i and 16 + i).
__device__ __forceinline__ void stash_row(const uint32_t (&a)[32], uint8_t* blk, int row, uint2* mask_out) {
  uint8_t* const dst = blk + stash_chunk_off(row, 0);
#pragma unroll
  for (int c = 0; c < 8; ++c)
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c * STASH_CHUNK_STRIDE), "r"(a[c * 4 + 0]),
                 "r"(a[c * 4 + 1]), "r"(a[c * 4 + 2]), "r"(a[c * 4 + 3])
                 : "memory");
  uint32_t mbits[2] = {0u, 0u};
#pragma unroll
  for (int w = 0; w < 32; ++w) {
    __nv_bfloat162 hv;
    *reinterpret_cast<uint32_t*>(&hv) = a[w];
    mbits[w >> 4] |= __hgt2_mask(hv, __nv_bfloat162(__float2bfloat16(0.f), __float2bfloat16(0.f))) & (0x00010001u << (w & 15));
  }
  *mask_out = make_uint2(mbits[0], mbits[1]);
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------- kernel
// TS = true (default): the hidden activations never touch shared memory on the MMA path.  The epilogue writes layer l's
// bf16 output with tcgen05.st over the first 128 columns of the accumulator it has just read, and layer l+1 takes its A
// operand from there (tcgen05.mma [d], [a], b-desc) while accumulating into the other buffer.  Per layer this removes the
// A-operand reads (64 of 192 KB) and, outside training, the activation writes (64 KB) from the shared-memory pipe.  The
// encodings (IPE, IDE) still arrive through shared memory.  TS = false (test build only, RSN_FWD_TS=0): the operand is
// written in place into the swizzled shared-memory activation blocks (bit-identical results).
// (The cta_group::2 CTA-pair form and the cluster-multicast weight stream of round 1 were validated bit for bit and
// lost -- 2.86 vs 2.63 ms and no change, DESIGN.md §4 -- and are no longer part of this kernel; their building blocks stay
// in umma.cuh and csrc/probe.cu.)
// TRAIN (needs TS): the launch carries four more warps that write the activation stash (see the header); p.stash != NULL.
template <bool TS, bool TRAIN>
__global__ void __launch_bounds__(TRAIN ? NUM_THREADS_TRAIN : NUM_THREADS, 1) field_fwd_kernel(const FwdParams p) {
  static_assert(TS || !TRAIN, "the training form takes its operands from TMEM");
  extern __shared__ __align__(1024) uint8_t smem[];   // no static shared memory in this kernel: the window starts here
  Barriers& bars = *reinterpret_cast<Barriers*>(smem + SMEM_BARS);
  constexpr bool SBIAS = true;
  // weight ring: 3 x 32 KB behind the shared-memory activation blocks (SS form); TS: no activation blocks -> 5 x 32 KB
  constexpr int NS = TS ? 5 : NUM_STAGES;
  constexpr int ring_off = TS ? SMEM_ACT : SMEM_W;
  constexpr uint32_t STB = W_STAGE_BYTES;
  constexpr uint32_t ARRIVALS = TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s_act = smem_u32(smem + SMEM_ACT);
  const uint32_t s_enc = smem_u32(smem + SMEM_ENC);
  const uint32_t s_w = smem_u32(smem + ring_off);
  const uint32_t s_bias = smem_u32(smem + SMEM_BIAS);
  if ((smem_u32(smem) & 1023u) != 0u) {   // the swizzled operand blocks need 1024-byte alignment
    if (threadIdx.x == 0) printf("rsn_b200: field_fwd_kernel: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  // valid points of this launch: the host's count, or rays counted on the device (bounce passes)
  const int n_points = (int)rsn_count((int64_t)p.n_points / p.n_samples, p.n_rays_dev) * p.n_samples;
  const int n_tiles = (n_points + TILE - 1) / TILE;

  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < MAX_STAGES; ++i) {
        mbar_init(&bars.w_full[i], 1);
        mbar_init(&bars.w_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars.enc_full[i], ARRIVALS);
        mbar_init(&bars.enc_empty[i], 1);
        mbar_init(&bars.acc_full[2 * i], 1);
        mbar_init(&bars.acc_full[2 * i + 1], 1);
        mbar_init(&bars.bias_full[i], 1);
      }
      for (int i = 0; i < 8; ++i) mbar_init(&bars.act_ready[i], ARRIVALS);
      for (int i = 0; i < 2; ++i) mbar_init(&bars.a_free[i], 4);
      mbar_init(&bars.ide_ready, ARRIVALS);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&bars.tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_slot;
  // tiles of this CTA: blockIdx.x + it * gridDim.x
  const int n_my_tiles = (n_tiles > (int)blockIdx.x) ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto tile_of = [&](int it) -> int { return (int)blockIdx.x + it * (int)gridDim.x; };
  auto arrive_issuer = [&](uint64_t* bar) { mbar_arrive(bar); };

  if (warp == 0) {
    // ===================================================================== weight producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_my_tiles; ++it) {
        uint32_t off = 0;
        for (int c = 0; c < N_FWD_CHUNKS; ++c) {
          const uint32_t bytes = fwd_chunk_bytes(c);
          mbar_wait(&bars.w_empty[stage], phase ^ 1);
          if (p.debug & 4) {
            mbar_arrive(&bars.w_full[stage]);     // timing experiment: no weight traffic at all (results are garbage)
          } else {
            mbar_expect_tx(&bars.w_full[stage], bytes);
            bulk_g2s(smem + ring_off + stage * (int)STB, p.wblob + off, bytes, &bars.w_full[stage]);
          }
          off += bytes;
          if (++stage == NS) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // The whole warp runs the control flow and the barrier waits; every single-thread instruction (tcgen05.mma,
    // tcgen05.commit, the bias bulk copies) is issued under elect_one_sync() by the warp's elected lane.
    {
      int stage = 0;
      uint32_t wphase = 0;
      uint32_t ar_phase = 0;  // bit 4 b + g: parity of the next completion of act_ready[4 b + g]
      int buf = 0;
      int use = 0;            // accumulator-buffer uses so far (use u accumulates into buffer u & 1)
      // TRAIN: before the first MMA of use u overwrites buffer u & 1, the stash warps must have read the operand that the
      // epilogue of use u - 2 left in its first 128 columns
      auto wait_a_free = [&]() {
        if (TRAIN && use >= 2) mbar_wait(&bars.a_free[use & 1], (uint32_t)((use - 2) >> 1) & 1u);
        ++use;
      };
      constexpr int MM = 128;
      constexpr uint32_t ID256 = instr_desc_bf16(MM, 256, 0, 0);
      constexpr uint32_t ID128 = instr_desc_bf16(MM, 128, 0, 0);
      constexpr uint32_t ID16 = instr_desc_bf16(MM, 16, 0, 0);
      constexpr uint32_t BDIV = 1;
      constexpr uint32_t HALF_B = 128u * 128u;   // bytes of 128 weight rows (one output half) inside a [256][64] K-block image
      auto commit = [&](uint64_t* bar) {
        if (elect_one_sync()) mma_commit(bar);
      };
      auto wait_in = [&](uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); };
      auto ring_wait = [&]() -> uint32_t {
        mbar_wait(&bars.w_full[stage], wphase);
        tc_fence_after();
        return s_w + (uint32_t)stage * STB;
      };
      auto ring_advance = [&]() -> int {     // -> the slot just consumed (to be released with ring_release_slot)
        const int s_ = stage;
        if (++stage == NS) {
          stage = 0;
          wphase ^= 1;
        }
        return s_;
      };
      auto ring_release_slot = [&](int s_) { commit(&bars.w_empty[s_]); };
      auto ring_release = [&]() { ring_release_slot(ring_advance()); };
      auto wait_act = [&](int g) {
        const int i = (buf ^ 1) * 4 + g;     // handed over by the previous use, which accumulated into the other buffer
        wait_in(&bars.act_ready[i], (ar_phase >> i) & 1u);
        ar_phase ^= (1u << i);
        tc_fence_after();
      };
      // K-major operands: LBO unused (16), SBO = 1024; one K=16 step advances both start addresses by 32 bytes
      constexpr uint32_t HI = desc_hi_sw128(1024);
      auto issue_kb = [&](uint32_t a_addr, uint32_t b_addr, int ksteps, uint32_t idesc, uint32_t tmem_d, bool& acc) {
        const uint32_t a_lo = desc_lo(a_addr, 16), b_lo = desc_lo(b_addr, 16);
        auto mma = [&](uint32_t k2, uint32_t accum) {
          mma_bf16_ss_lo(tmem_d, a_lo + k2, b_lo + k2, HI, idesc, accum);
        };
        if (elect_one_sync()) {
          mma(0, acc ? 1u : 0u);
          mma(2, 1u);
          mma(4, 1u);
          if (ksteps == 4) mma(6, 1u);
        }
        acc = true;
      };
      // K-block g of the hidden activations: shared-memory block g (SS) or TMEM columns [32 g, 32 g + 32) of the buffer
      // the previous layer accumulated into (TS)
      auto issue_act = [&](int g, uint32_t b_addr, uint32_t idesc, uint32_t tmem_d, uint32_t a_tm, bool& acc) {
        if (!TS) {
          issue_kb(s_act + g * BLOCK_BYTES, b_addr, 4, idesc, tmem_d, acc);
        } else {
          const uint32_t b_lo = desc_lo(b_addr, 16), a0 = a_tm + (uint32_t)g * 32u;
          if (elect_one_sync()) {
            mma_bf16_ts_lo(tmem_d, a0, b_lo, HI, idesc, acc ? 1u : 0u);
            mma_bf16_ts_lo(tmem_d, a0 + 8, b_lo + 2, HI, idesc, 1u);
            mma_bf16_ts_lo(tmem_d, a0 + 16, b_lo + 4, HI, idesc, 1u);
            mma_bf16_ts_lo(tmem_d, a0 + 24, b_lo + 6, HI, idesc, 1u);
          }
          acc = true;
        }
      };
      // Bias of wide layer j (10 per tile) -> slot j & 1, requested as soon as the epilogue of layer j - 2 has published
      // its last group (the issuer sees that as the last act_ready wait of layer j - 1).
      auto request_bias = [&](int j) {
        if (j >= n_my_tiles * WIDE_LAYERS || !elect_one_sync()) return;
        const int jl = j % WIDE_LAYERS;
        uint8_t* const slot = smem + SMEM_BIAS + (j & 1) * BIAS_SLOT_BYTES;
        if (jl < 8) {
          mbar_expect_tx(&bars.bias_full[j & 1], 1024u);
          bulk_g2s(slot, p.bias + BIAS_BASE + jl * 256, 1024u, &bars.bias_full[j & 1]);
        } else if (jl == 8) {   // bottleneck (256) + the 16 head biases behind it in the vector -> [0, 1088)
          static_assert(BIAS_HEAD == BIAS_BOTT + 256, "head biases follow the bottleneck bias");
          mbar_expect_tx(&bars.bias_full[j & 1], 1088u);
          bulk_g2s(slot, p.bias + BIAS_BOTT, 1088u, &bars.bias_full[j & 1]);
        } else {                // mid (128) -> [0, 512), rgb (16) -> [1088, 1152): the slot's wide part is refilled before the
                                // rgb epilogue runs, its tail only by the next mid layer
          mbar_expect_tx(&bars.bias_full[j & 1], 512u + 64u);
          bulk_g2s(slot, p.bias + BIAS_MID, 512u, &bars.bias_full[j & 1]);
          bulk_g2s(slot + 1088, p.bias + BIAS_RGB, 64u, &bars.bias_full[j & 1]);
        }
      };
      request_bias(0);
      request_bias(1);
      RSN_TRACE_DECL(0);
      for (int it = 0; it < n_my_tiles; ++it) {
        const int eb = it & 1;
        const uint32_t enc_a = s_enc + (uint32_t)eb * 2 * BLOCK_BYTES;
        bool acc;
        const bool tr = (p.debug & 32) && blockIdx.x == 0 && it == 2 && lane == 0;
        // ---- base layers 0..7
        for (int l = 0; l < 8; ++l) {
          const uint32_t tm = tmem + (uint32_t)buf * 256;
          acc = false;
          wait_a_free();
          if (l == 0) {
            wait_in(&bars.enc_full[eb], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
          }
          // The 256 outputs of a wide layer are two N = 128 accumulations (columns 0-127 | 128-255 of the buffer, weight rows
          // 0-127 | 128-255 of every K-block image).  Same tensor time as N = 256, but the first half is COMMITTED while the
          // tensor pipe still works on the second: the epilogue converts groups 0, 1 (and the next layer's first K-blocks
          // start) two K-block times earlier -- with one tile in flight, the stretch between the last hand-over of layer l
          // and the first of layer l + 1 is where the tensor pipe idled (in-kernel trace: ~1,100 of ~3,200 cycles per layer).
          // Issue order: [kb0: h0 h1] [kb1: h0 h1] [kb2: h0] [kb3: h0] commit(h0) [kb2: h1] [kb3: h1] commit(h1).
          bool acc1 = false;                                  // (acc = half 0, acc1 = half 1)
          const uint32_t a_tm = tmem + (uint32_t)(buf ^ 1) * 256;
          if (l == 0 || l == 4) {
            uint32_t w = ring_wait();
            issue_kb(enc_a, w, 4, ID128, tm, acc);
            issue_kb(enc_a, w + HALF_B, 4, ID128, tm + 128, acc1);
            ring_release();
            w = ring_wait();
            issue_kb(enc_a + BLOCK_BYTES, w, ENC_KSTEPS_B1, ID128, tm, acc);
            if (l == 0) commit(&bars.acc_full[2 * buf]);
            issue_kb(enc_a + BLOCK_BYTES, w + HALF_B, ENC_KSTEPS_B1, ID128, tm + 128, acc1);
            ring_release();
            if (l == 0) commit(&bars.acc_full[2 * buf + 1]);
          }
          if (l > 0 && TRAIN && !RSN_FWD_TRAIN_SPLIT) {
            for (int g = 0; g < 4; ++g) {
              wait_act(g);
              if (g == 3) request_bias(it * WIDE_LAYERS + l + 1);
              const uint32_t w = ring_wait();
              issue_act(g, w, ID256, tm, a_tm, acc);
              ring_release();
            }
            commit(&bars.acc_full[2 * buf]);
            commit(&bars.acc_full[2 * buf + 1]);
          } else if (l > 0) {
            for (int g = 0; g < 2; ++g) {
              wait_act(g);
              RSN_TRACE(tr, 1000 + 10 * l + g);
              const uint32_t w = ring_wait();
              issue_act(g, w, ID128, tm, a_tm, acc);
              issue_act(g, w + HALF_B, ID128, tm + 128, a_tm, acc1);
              ring_release();
            }
            wait_act(2);
            RSN_TRACE(tr, 1000 + 10 * l + 2);
            const uint32_t w2 = ring_wait();
            const int s2 = ring_advance();
            issue_act(2, w2, ID128, tm, a_tm, acc);
            wait_act(3);
            RSN_TRACE(tr, 1000 + 10 * l + 3);
            request_bias(it * WIDE_LAYERS + l + 1);
            const uint32_t w3 = ring_wait();
            const int s3 = ring_advance();
            issue_act(3, w3, ID128, tm, a_tm, acc);
            commit(&bars.acc_full[2 * buf]);
            issue_act(2, w2 + HALF_B, ID128, tm + 128, a_tm, acc1);
            ring_release_slot(s2);
            issue_act(3, w3 + HALF_B, ID128, tm + 128, a_tm, acc1);
            ring_release_slot(s3);
            commit(&bars.acc_full[2 * buf + 1]);
          }
          RSN_TRACE(tr, 1500 + l);
          buf ^= 1;
        }
        // ---- layer 8: bottleneck (N=256) + heads (N=16, other buffer, columns 240..255)
        {
          const uint32_t tm = tmem + (uint32_t)buf * 256;
          acc = false;
          wait_a_free();
          const uint32_t a_tm = tmem + (uint32_t)(buf ^ 1) * 256;
          bool acc1 = false;
          for (int g = 0; g < 2; ++g) {
            wait_act(g);
            const uint32_t w = ring_wait();
            issue_act(g, w, ID128, tm, a_tm, acc);
            issue_act(g, w + HALF_B, ID128, tm + 128, a_tm, acc1);
            ring_release();
          }
          wait_act(2);
          const uint32_t w2 = ring_wait();
          const int s2 = ring_advance();
          issue_act(2, w2, ID128, tm, a_tm, acc);
          wait_act(3);
          request_bias(it * WIDE_LAYERS + 9);
          const uint32_t w3 = ring_wait();
          const int s3 = ring_advance();
          issue_act(3, w3, ID128, tm, a_tm, acc);
          // The heads (N = 16, all four K-blocks of h7).  TS form: they ride behind the second half, whose commit covers them
          // (the epilogue reads them after groups 2, 3).  SS form: the epilogue overwrites the shared-memory blocks 0, 1 in
          // place as soon as the FIRST half is committed, so every MMA that reads them -- the heads too -- goes before it.
          auto issue_heads = [&](uint32_t w) {
            bool acc_h = false;
            for (int kb = 0; kb < 4; ++kb)
              issue_act(kb, w + kb * (N_HEAD * 128 / BDIV), ID16, tmem + (uint32_t)(buf ^ 1) * 256 + HEAD_TMEM_COL, a_tm,
                        acc_h);
          };
          int sh = -1;
          if (!TS) {
            const uint32_t w = ring_wait();
            sh = ring_advance();
            issue_heads(w);
          }
          commit(&bars.acc_full[2 * buf]);
          issue_act(2, w2 + HALF_B, ID128, tm + 128, a_tm, acc1);
          ring_release_slot(s2);
          issue_act(3, w3 + HALF_B, ID128, tm + 128, a_tm, acc1);
          ring_release_slot(s3);
          if (TS) {
            const uint32_t w = ring_wait();
            issue_heads(w);
            ring_release();
          } else {
            ring_release_slot(sh);
          }
          commit(&bars.acc_full[2 * buf + 1]);
          buf ^= 1;
        }
        // ---- layer 9: mid MLP, A = [bottleneck 256 | IDE 48], N = 128
        {
          const uint32_t tm = tmem + (uint32_t)buf * 256;
          acc = false;
          wait_a_free();
          for (int c = 0; c < 2; ++c) {
            wait_act(2 * c);
            wait_act(2 * c + 1);
            if (c == 1) request_bias(it * WIDE_LAYERS + 10);
            const uint32_t w = ring_wait();
            issue_act(2 * c, w, ID128, tm, tmem + (uint32_t)(buf ^ 1) * 256, acc);
            issue_act(2 * c + 1, w + 128 * 128 / BDIV, ID128, tm, tmem + (uint32_t)(buf ^ 1) * 256, acc);
            ring_release();
          }
          wait_in(&bars.ide_ready, (uint32_t)it & 1u);
          tc_fence_after();
          const uint32_t w = ring_wait();
          issue_kb(enc_a, w, IDE_KSTEPS, ID128, tm, acc);
          ring_release();
          commit(&bars.acc_full[2 * buf]);
          commit(&bars.enc_empty[eb]);
          buf ^= 1;
        }
        // ---- layer 10: rgb head, A = mid hidden 128, N = 16
        {
          const uint32_t tm = tmem + (uint32_t)buf * 256;
          acc = false;
          wait_a_free();
          wait_act(0);
          wait_act(1);
          request_bias(it * WIDE_LAYERS + 11);
          const uint32_t w = ring_wait();
          issue_act(0, w, ID16, tm, tmem + (uint32_t)(buf ^ 1) * 256, acc);
          issue_act(1, w + N_HEAD * 128 / BDIV, ID16, tm, tmem + (uint32_t)(buf ^ 1) * 256, acc);
          ring_release();
          commit(&bars.acc_full[2 * buf]);
          buf ^= 1;
        }
        RSN_TRACE_END(tr, 0);
      }
    }
  } else if (warp >= 6 && warp < 10) {
    // ===================================================================== epilogue warps (thread = point row)
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t af_phase = 0;
    int buf = 0;
    int jw = 0;   // wide layers converted so far
    RSN_TRACE_DECL(1);
    // bias slot of the wide layer about to be converted (waits until its bulk copy has landed)
    auto bias_slot = [&]() -> uint32_t {
      mbar_wait(&bars.bias_full[jw & 1], (uint32_t)(jw >> 1) & 1u);
      const uint32_t a = s_bias + (uint32_t)(jw & 1) * BIAS_SLOT_BYTES;
      ++jw;
      return a;
    };
    for (int it = 0; it < n_my_tiles; ++it) {
      const int eb = it & 1;
      const int tile = tile_of(it);
      const int pt = tile * TILE + row;
      const bool valid = pt < n_points;
      uint8_t* const st = TRAIN ? p.stash + (size_t)tile * STASH_TILE_BYTES : nullptr;
      auto wait_acc = [&](int h = 0) {     // output half h of the layer accumulating into `buf` is complete
        const int i = 2 * buf + h;
        mbar_wait(&bars.acc_full[i], (af_phase >> i) & 1u);
        af_phase ^= (1u << i);
        tc_fence_after();
      };
      // One 64-column group of a wide layer: convert and hand to the issuer (and, in training, to the stash warps):
      // act_ready[buf][g]
      auto convert = [&](auto relu_c, uint32_t sb, int g) {
        constexpr bool RELU = decltype(relu_c)::value;
        const uint32_t acc_c = tlane + (uint32_t)buf * 256 + g * 64, a_t = tlane + (uint32_t)buf * 256 + g * 32;
        epilogue_group<RELU, TS>(acc_c, s_act + g * BLOCK_BYTES, row, sb + g * 256, a_t);
        if (TS) tmem_st_wait(); else fence_proxy_async();
        tc_fence_before();
        arrive_issuer(&bars.act_ready[buf * 4 + g]);
        RSN_TRACE((p.debug & 32) && blockIdx.x == 0 && it == 2 && warp == 6 && lane == 0, 2100 + 10 * ((jw - 1) % WIDE_LAYERS) + g);
      };
      using T_ = std::integral_constant<bool, true>;
      using F_ = std::integral_constant<bool, false>;
      const bool tr = (p.debug & 32) && blockIdx.x == 0 && it == 2 && warp == 6 && lane == 0;
      for (int l = 0; l < 8; ++l) {
        wait_acc();
        RSN_TRACE(tr, 2000 + l);
        const uint32_t sb = bias_slot();
        for (int g = 0; g < 4; ++g) {
          if (g == 2) wait_acc(1);
          convert(T_{}, sb, g);
        }
        buf ^= 1;
      }
      // ---- layer 8: bottleneck -> activation blocks (no activation), then heads + IDE
      float diff[3], tint[3];
      {
        wait_acc();
        const uint32_t sb = bias_slot();
        // (not stashed: the wgrad derives everything that involves the bottleneck from h7, csrc/field_wgrad.cu)
        for (int g = 0; g < 4; ++g) {
          if (g == 2) wait_acc(1);
          convert(F_{}, sb, g);
        }
        uint32_t hv[16];
        tmem_ld16(tlane + (uint32_t)(buf ^ 1) * 256 + HEAD_TMEM_COL, hv);
        float hb[12];   // head biases: behind the bottleneck bias in this layer's slot (only this layer's request writes there)
#pragma unroll
        for (int i = 0; i < 3; ++i)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(hb[i * 4 + 0]), "=f"(hb[i * 4 + 1]), "=f"(hb[i * 4 + 2]), "=f"(hb[i * 4 + 3])
                       : "r"(sb + 1024u + (uint32_t)i * 16u));
        tmem_ld_wait();
        float h[11];
#pragma unroll
        for (int i = 0; i < 11; ++i) h[i] = __uint_as_float(hv[i]) + hb[i];
        // view direction of this point's ray
        float d[3] = {0.f, 0.f, 1.f};
        if (valid) {
          const int ray = pt / p.n_samples;
#pragma unroll
          for (int a = 0; a < 3; ++a) d[a] = __ldg(p.dirs + (size_t)ray * 3 + a);
        }
        const float raw_sigma = h[0];
        const float sigma = softplusf(raw_sigma + 0.5f);  // density_bias = 0.5 (field.py:46,136)
        // pred normals = normalize(-normalize(Linear(emb)))  (field.py:139-144)
        float n[3] = {h[1], h[2], h[3]};
        float nn = fmaxf(sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), 1e-12f);
#pragma unroll
        for (int a = 0; a < 3; ++a) n[a] = -(n[a] / nn);
        nn = fmaxf(sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), 1e-12f);
#pragma unroll
        for (int a = 0; a < 3; ++a) n[a] = n[a] / nn;
        const float ndd = d[0] * n[0] + d[1] * n[1] + d[2] * n[2];  // field.py:204
        const float rough_sp = softplusf(h[4]);                     // model.py:173 (Softplus -> IDE)
        const float rough_sg = sigmoid_acc(h[4]);                   // model.py:225 (Sigmoid -> rendered)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          diff[a] = sigmoid_acc(h[5 + a]);
          tint[a] = sigmoid_acc(h[8 + a]);
        }
        // IDE of the VIEW direction (App. B Q1) -> first 48 columns of this tile's enc block 0
        float t[48];
        if (p.mode != 1) {
          ide_features(d, (p.pt_rho && valid) ? __ldg(p.pt_rho + pt) : rough_sp, t);
        } else {
#pragma unroll
          for (int i = 0; i < 48; ++i) t[i] = 0.f;  // get_inf_color feeds a zero IDE (field.py:199)
        }
        const uint32_t ide_blk = s_enc + (uint32_t)eb * 2 * BLOCK_BYTES;
        const uint32_t ide_row = ide_blk + (uint32_t)row * 128u;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const float f8[8] = {t[c * 8 + 0], t[c * 8 + 1], t[c * 8 + 2], t[c * 8 + 3],
                               t[c * 8 + 4], t[c * 8 + 5], t[c * 8 + 6], t[c * 8 + 7]};
          store_chunk(ide_row, row, c, f8);
        }
        if (st) {   // columns 48..63 are not read by the mid MMA but the wgrad reads the whole block
          const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          store_chunk(ide_row, row, 6, z8);
          store_chunk(ide_row, row, 7, z8);
        }
        // publish the IDE block (rows of this warp): in training, bulk-store the slice first and wait until the TMA engine has
        // read it (the prologue warps overwrite the block with the next-but-one tile's encoding)
        if (TRAIN) {
          warp_store_rows(st + (size_t)STASH_IDE * BLOCK_BYTES, ide_blk, q, lane);
          warp_store_guard<0>(lane);
        } else {
          fence_proxy_async();
        }
        tc_fence_before();
        arrive_issuer(&bars.ide_ready);
        if (valid && p.aux) {
          p.aux[(size_t)pt * 8 + 3] = h[1];
          p.aux[(size_t)pt * 8 + 4] = h[2];
          p.aux[(size_t)pt * 8 + 5] = h[3];
          p.aux[(size_t)pt * 8 + 6] = h[4];   // raw roughness head (Field.get_roughness with an arbitrary activation)
        }
        if (valid) {
          p.sigma[pt] = sigma;
          float4* fo = reinterpret_cast<float4*>(p.feat + (size_t)pt * N_FEAT);
          fo[1] = make_float4(diff[1], diff[2], tint[0], tint[1]);           // cols 4..7
          fo[2] = make_float4(tint[2], n[0], n[1], n[2]);                    // cols 8..11
          fo[3] = make_float4(rough_sg, ndd, raw_sigma, rough_sp);           // cols 12..15
        }
        buf ^= 1;
      }
      // ---- layer 9: mid hidden (ReLU) -> activation blocks 0,1
      {
        wait_acc();
        const uint32_t sb = bias_slot();
        for (int g = 0; g < 2; ++g) convert(T_{}, sb, g);
        buf ^= 1;
      }
      // ---- layer 10: rgb
      {
        wait_acc();
        uint32_t rv[16];
        tmem_ld16(tlane + (uint32_t)buf * 256, rv);
        float4 rb;      // rgb bias: tail of the mid layer's slot (slot 1 of every tile: WIDE_LAYERS is even)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(rb.x), "=f"(rb.y), "=f"(rb.z), "=f"(rb.w)
                     : "r"(s_bias + BIAS_SLOT_BYTES + 1088u));
        tmem_ld_wait();
        tc_fence_before();
        const float mid[3] = {sigmoid_acc(__uint_as_float(rv[0]) + rb.x), sigmoid_acc(__uint_as_float(rv[1]) + rb.y),
                              sigmoid_acc(__uint_as_float(rv[2]) + rb.z)};
        if (valid) {
          float rgb[3];
#pragma unroll
          for (int a = 0; a < 3; ++a) rgb[a] = (p.mode != 1) ? diff[a] + tint[a] * mid[a] : mid[a];
          float4* fo = reinterpret_cast<float4*>(p.feat + (size_t)pt * N_FEAT);
          fo[0] = make_float4(rgb[0], rgb[1], rgb[2], diff[0]);  // cols 0..3
          if (p.aux) {
            p.aux[(size_t)pt * 8 + 0] = mid[0];
            p.aux[(size_t)pt * 8 + 1] = mid[1];
            p.aux[(size_t)pt * 8 + 2] = mid[2];
          }
        }
        buf ^= 1;
      }
      RSN_TRACE_END(tr, 1);
    }
  } else if (warp >= 10) {
    // ===================================================================== stash warps (training launches only)
    if constexpr (TRAIN) {
      const int q = warp & 3;                                   // TMEM lane quarter this warp may touch
      const int row = q * 32 + lane;
      const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
      uint32_t ar_phase = 0;                                    // bit 4 b + g: parity of the next completion of act_ready[4 b + g]
      int sbuf = 0;                                             // accumulator buffer of the layer use being stashed
      auto wait_group = [&](int g) {
        const int i = sbuf * 4 + g;
        mbar_wait(&bars.act_ready[i], (ar_phase >> i) & 1u);
        ar_phase ^= (1u << i);
        tc_fence_after();
      };
      auto release = [&]() {                                    // this warp no longer reads the operand in buffer sbuf
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.a_free[sbuf]);
        sbuf ^= 1;
      };
      for (int it = 0; it < n_my_tiles; ++it) {
        uint8_t* const st = p.stash + (size_t)tile_of(it) * STASH_TILE_BYTES;
        uint2* const masks = reinterpret_cast<uint2*>(st + STASH_MASK_OFF);
        // one layer use with NG stashed groups: mask layer ml, stash blocks blk0 .. blk0 + NG - 1
        // RSN_STASH_LAG_FWD = k: group g is read and stored once the epilogue has handed group min(g + k, last) over
        // (0: right behind its own hand-over; 4: the whole layer after its last hand-over, i.e. in the stretch in which the
        // epilogue warps of the same SMSPs wait for the next layer's accumulator).  The operand stays valid until the issuer
        // re-uses the buffer two layers later (a_free).  Measured over 30 sustained launches, alternating builds: with the
        // two-halves commit (RSN_FWD_TRAIN_SPLIT, which shortens that stretch) k = 0 / 1 / 2 / 3: 3.07 / 3.09 / 2.98-3.05 /
        // 3.2 ms; with plain N = 256 layers k = 0 / 2 / 3 / 4: 3.06 / 2.95 / 2.94 / 2.91 ms against 3.01 ms for the previous
        // default (split, k = 0).  Training therefore runs N = 256 layers with k = 4; inference keeps the split.
        auto stash_layer = [&](int ng, int ml, int blk0) {
          int waited = 0;
          for (int g = 0; g < ng; ++g) {
            const int need = min(g + RSN_STASH_LAG_FWD, ng - 1);
            while (waited <= need) wait_group(waited++);
            uint32_t a[32];
            tmem_ld32(tlane + (uint32_t)sbuf * 256 + (uint32_t)g * 32u, a);
            tmem_ld_wait();
            if (g == ng - 1) release();
            stash_row(a, st + (size_t)(blk0 + g) * BLOCK_BYTES, row, masks + mask_entry(ml, g, row));
          }
        };
        for (int l = 0; l < 8; ++l) stash_layer(4, l, STASH_H + 4 * l);
        for (int g = 0; g < 4; ++g) wait_group(g);              // bottleneck: not stashed (csrc/field_wgrad.cu), only observed
        release();
        stash_layer(2, 8, STASH_MIDH);                          // mid hidden
        release();                                              // rgb: hands nothing over; one release per layer use
      }
    }
  } else {
    // ===================================================================== prologue warps (next tile's IPE)
    const int row = (warp - 2) * 32 + lane;
    for (int it = 0; it < n_my_tiles; ++it) {
      const int eb = it & 1;
      const int tile = tile_of(it);
      const int pt = tile * TILE + row;
      const bool stash_on = TRAIN;
      float xm[3] = {0.f, 0.f, 0.f}, dg[3] = {0.f, 0.f, 0.f};
      if (pt < n_points) {
        if (p.mode == 0) {
          const int ray = pt / p.n_samples;
          const int s = pt - ray * p.n_samples;
          float o[3], d[3];
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            o[a] = __ldg(p.origins + (size_t)ray * 3 + a);
            d[a] = __ldg(p.dirs + (size_t)ray * 3 + a);
          }
          const float* b = p.bins + (size_t)ray * (p.n_samples + 1) + s;
          frustum_gaussian_contracted(o, d, __ldg(b), __ldg(b + 1), __ldg(p.area + ray), xm, dg);
        } else if (p.mode == 2) {
          // Field.get_density(mean, cov) on caller-supplied Gaussians (field.py:122-137): the encoding reads diag(cov) only
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            xm[a] = __ldg(p.pt_mean + (size_t)pt * 3 + a);
            dg[a] = __ldg(p.pt_cov + (size_t)pt * 9 + a * 4);
          }
        } else {
          // get_inf_color (field.py:190-201): mean = 2 w, cov = 0.6 sqradius (I - w w^T), NOT contracted
          const float sq = __fmul_rn(0.6f, __ldg(p.area + pt));
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const float w = __ldg(p.dirs + (size_t)pt * 3 + a);
            xm[a] = __fmul_rn(2.0f, w);
            dg[a] = __fmul_rn(sq, __fsub_rn(1.0f, __fmul_rn(w, w)));
          }
        }
      }
      if (p.debug & 2) xm[0] = xm[1] = xm[2] = 0.f, dg[0] = dg[1] = dg[2] = 100.f;
      mbar_wait(&bars.enc_empty[eb], ((uint32_t)(it >> 1) & 1u) ^ 1u);
      const uint32_t enc_blk = s_enc + (uint32_t)eb * 2 * BLOCK_BYTES;
      encode_row(enc_blk, row, xm, dg, stash_on);
      if (stash_on) {
        // stash the two enc blocks (this warp's rows) and wait until the TMA engine has READ them: the epilogue
        // warps overwrite block 0 with the IDE later in the tile
        uint8_t* se = p.stash + (size_t)tile * STASH_TILE_BYTES + (size_t)STASH_ENC * BLOCK_BYTES;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          bulk_s2g_u32(se + (warp - 2) * 4096, enc_blk + (uint32_t)(warp - 2) * 4096u, 4096);
          bulk_s2g_u32(se + BLOCK_BYTES + (warp - 2) * 4096, enc_blk + BLOCK_BYTES + (uint32_t)(warp - 2) * 4096u, 4096);
          bulk_commit();
          bulk_wait_read<0>();
        }
        __syncwarp();
      } else {
        fence_proxy_async();
      }
      arrive_issuer(&bars.enc_full[eb]);
    }
  }

  if (TRAIN && lane == 0 && warp >= 2 && warp < 10) bulk_wait_all<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

extern "C" int rsn_ipe_freqs(float* host_out16) {
  RSN_ARG(host_out16 != nullptr, "rsn_ipe_freqs: null pointer");
  for (int i = 0; i < 16; ++i) host_out16[i] = h_freq[i];
  return 0;
}

extern "C" int64_t rsn_field_blob_bytes(void) { return (int64_t)FWD_BLOB_BYTES; }
extern "C" int64_t rsn_field_bias_count(void) { return (int64_t)N_BIAS; }

extern "C" int64_t rsn_field_stash_bytes(int64_t n_points) {
  return ((n_points + TILE - 1) / TILE) * (int64_t)STASH_TILE_BYTES;
}

namespace {
std::atomic<unsigned long long> g_fwd_smem_done[3];

int launch_fwd(FwdParams& p, cudaStream_t stream) {
  p.n_tiles = (p.n_points + TILE - 1) / TILE;
  p.debug = rsn_env_int("RSN_FWD_DEBUG", 0);
  const size_t smem = SMEM_TOTAL;
  const int grid = std::min(p.n_tiles, rsn_num_sms());
  if (p.stash) {   // training: four more warps write the activation stash
    RSN_CUDA(rsn_ensure_smem(field_fwd_kernel<true, true>, (int)smem, g_fwd_smem_done[0]));
    field_fwd_kernel<true, true><<<grid, NUM_THREADS_TRAIN, smem, stream>>>(p);
    RSN_LAUNCH_CHECK("field_fwd_kernel");
    return 0;
  }
#ifdef RSN_DEBUG_SWITCHES
  // RSN_FWD_TS=0 selects the shared-memory (SS) operand form (same results bit for bit; test build, inference only)
  if (rsn_env_int("RSN_FWD_TS", 1) == 0) {
    RSN_CUDA(rsn_ensure_smem(field_fwd_kernel<false, false>, (int)smem, g_fwd_smem_done[2]));
    field_fwd_kernel<false, false><<<grid, NUM_THREADS, smem, stream>>>(p);
    RSN_LAUNCH_CHECK("field_fwd_kernel");
    return 0;
  }
#endif
  RSN_CUDA(rsn_ensure_smem(field_fwd_kernel<true, false>, (int)smem, g_fwd_smem_done[1]));
  field_fwd_kernel<true, false><<<grid, NUM_THREADS, smem, stream>>>(p);
  RSN_LAUNCH_CHECK("field_fwd_kernel");
  return 0;
}
}  // namespace

extern "C" int rsn_field_forward_train(const void* wblob, const float* bias, int mode, const float* origins,
                                       const float* dirs, const float* area, const float* bins, int64_t n_rays,
                                       int64_t n_samples, float* sigma, float* feat, void* stash, float* aux,
                                       const int* n_rays_dev, cudaStream_t stream) {
  RSN_ARG(mode == 0 || mode == 1, "rsn_field_forward: mode must be 0 (samples) or 1 (infinity colour)");
  RSN_ARG(n_rays >= 0 && n_samples >= 1, "rsn_field_forward: bad shape");
  RSN_ARG(mode == 0 || n_samples == 1, "rsn_field_forward: mode 1 takes one point per ray");
  if (n_rays == 0) return 0;
  RSN_ARG(n_rays * n_samples < (int64_t)2147483647 - TILE, "rsn_field_forward: more than 2^31 points in one call");
  RSN_ARG(wblob && bias && dirs && area && sigma && feat, "rsn_field_forward: null pointer");
  RSN_ARG(mode == 1 || (origins && bins), "rsn_field_forward: origins/bins required in mode 0");
  RSN_ARG(((uintptr_t)wblob & 15) == 0 && ((uintptr_t)bias & 15) == 0 && ((uintptr_t)feat & 15) == 0,
          "rsn_field_forward: wblob/bias/feat must be 16-byte aligned");
  RSN_ARG(((uintptr_t)stash & 15) == 0, "rsn_field_forward: stash must be 16-byte aligned");
  FwdParams p = {};
  p.wblob = (const uint8_t*)wblob;
  p.bias = bias;
  p.mode = mode;
  p.origins = origins;
  p.dirs = dirs;
  p.area = area;
  p.bins = bins;
  p.n_samples = (int)n_samples;
  p.n_points = (int)(n_rays * n_samples);
  p.n_rays_dev = n_rays_dev;
  p.sigma = sigma;
  p.feat = feat;
  p.stash = (uint8_t*)stash;
  p.aux = aux;
  return launch_fwd(p, stream);
}

extern "C" int rsn_field_forward(const void* wblob, const float* bias, int mode, const float* origins,
                                 const float* dirs, const float* area, const float* bins, int64_t n_rays,
                                 int64_t n_samples, float* sigma, float* feat, const int* n_rays_dev,
                                 cudaStream_t stream) {
  return rsn_field_forward_train(wblob, bias, mode, origins, dirs, area, bins, n_rays, n_samples, sigma, feat,
                                 nullptr, nullptr, n_rays_dev, stream);
}

extern "C" int rsn_field_forward_points(const void* wblob, const float* bias, const float* mean, const float* cov,
                                        const float* dirs, const float* rho_override, int64_t n_points, float* sigma,
                                        float* feat, float* aux, cudaStream_t stream) {
  RSN_ARG(n_points >= 0, "rsn_field_forward_points: bad shape");
  if (n_points == 0) return 0;
  RSN_ARG(n_points < (int64_t)2147483647 - TILE, "rsn_field_forward_points: more than 2^31 points in one call");
  RSN_ARG(wblob && bias && mean && cov && dirs && sigma && feat, "rsn_field_forward_points: null pointer");
  RSN_ARG(((uintptr_t)wblob & 15) == 0 && ((uintptr_t)bias & 15) == 0 && ((uintptr_t)feat & 15) == 0,
          "rsn_field_forward_points: wblob/bias/feat must be 16-byte aligned");
  FwdParams p = {};
  p.wblob = (const uint8_t*)wblob;
  p.bias = bias;
  p.mode = 2;
  p.dirs = dirs;
  p.pt_mean = mean;
  p.pt_cov = cov;
  p.pt_rho = rho_override;
  p.n_samples = 1;
  p.n_points = (int)n_points;
  p.sigma = sigma;
  p.feat = feat;
  p.aux = aux;
  return launch_fwd(p, stream);
}

#ifdef RSN_DEBUG_SWITCHES
// Test build: copies the (tag, clock) trace of the last traced launch to host_out[2 * max_pairs]; returns the pair count
// and resets the trace.  Synchronises the device.
extern "C" int rsn_debug_fwd_trace(long long* host_out, int max_pairs) {
  int n[2] = {0, 0};
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  cudaMemcpyFromSymbol(n, g_fwd_trace_n, sizeof(n));
  int total = 0;
  for (int r = 0; r < 2; ++r) {
    const int k = std::min(std::min(n[r], 2047), max_pairs - total);
    if (k > 0) cudaMemcpyFromSymbol(host_out + 2 * total, g_fwd_trace, (size_t)k * 2 * sizeof(long long), (size_t)r * 4096 * sizeof(long long));
    total += std::max(k, 0);
  }
  const int zero[2] = {0, 0};
  cudaMemcpyToSymbol(g_fwd_trace_n, zero, sizeof(zero));
  return total;
}
#endif
