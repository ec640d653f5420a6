// Body of the dgrad-chain kernels (see field_bwd.cu for the description).
#pragma once
#include "rsn_common.cuh"
#include "umma.cuh"
#include "field_layout.cuh"
#include <algorithm>
#include <type_traits>
#include <stdlib.h>

namespace {

using namespace umma;
using namespace rsnf;

#ifndef RSN_BWD_SPLIT
#define RSN_BWD_SPLIT 0
#endif
#ifndef RSN_STASH_LAG_BWD
#define RSN_STASH_LAG_BWD 4
#endif
constexpr int KIND_NORMALS = 0, KIND_BACKWARD = 1;
constexpr int B_THREADS = 224;                    // warps 0 weights, 1 issuer, 2-5 epilogue, 6 stashed-encoding producer
constexpr int B_THREADS_STASH = 352;              // BACKWARD (TS form): + warps 7-10, which write the dY stash
constexpr int NUM_WSTAGES = 3, NUM_MSTAGES = 2;   // the second ring only carries the two stashed enc blocks
constexpr int MAX_WSTAGES = 5;                    // TS form: no shared-memory activation blocks -> 5 stages
// SS form: [4 dY blocks, rewritten in place step by step][seed block][3-stage weight ring][stashed-encoding ring]
// TS form: [5-stage weight ring][seed block (BACKWARD)][stashed-encoding ring]
constexpr int SM_ACT = 0;
constexpr int SM_SEED = 4 * BLOCK_BYTES;                   // 1 block: d(rgb head) cols 0-15, d(heads) cols 16-31
constexpr int SM_W = SM_SEED + BLOCK_BYTES;                // weight ring
constexpr int SM_M = SM_W + NUM_WSTAGES * W_STAGE_BYTES;   // stashed-encoding ring (IPE Jacobians)
constexpr int SM_TOTAL = SM_M + NUM_MSTAGES * BLOCK_BYTES; // 212,992
constexpr int SM_SEED_TS = MAX_WSTAGES * W_STAGE_BYTES;    // TS form: the seed block sits behind the 5-stage ring
static_assert(SM_SEED_TS + BLOCK_BYTES <= SM_M, "TS layout: ring + seed block overlap the encoding ring");

__constant__ float c_freq_b[16] = {
    0x1.0000000000000p+0f,  0x1.0c1b780000000p+1f,  0x1.18c9880000000p+2f,  0x1.26111c0000000p+3f,
    0x1.33f9760000000p+4f,  0x1.428a320000000p+5f,  0x1.51cb4e0000000p+6f,  0x1.61c5140000000p+7f,
    0x1.7280340000000p+8f,  0x1.8405f60000000p+9f,  0x1.965fde0000000p+10f, 0x1.a998080000000p+11f,
    0x1.bdb8d20000000p+12f, 0x1.d2cd4c0000000p+13f, 0x1.e8e1020000000p+14f, 0x1.0000000000000p+16f};

struct BwdParams {
  const uint8_t* wblob_t;   // transposed weight blob (BWD_BLOB_BYTES)
  const uint8_t* x_stash;   // forward stash [n_tiles][STASH_BLOCKS][16 KB]
  int kind;                 // KIND_NORMALS | KIND_BACKWARD
  int mode;                 // 0 = frustum samples, 1 = infinity colour
  int want_area;            // BACKWARD: continue to d pixel_area (mode 0) / d sqradius (mode 1)
  const float* origins;     // [N,3]
  const float* dirs;        // [N,3]
  const float* area;        // [N]
  const float* bins;        // [N,S+1]
  int n_samples, n_points, n_tiles;   // n_points / n_tiles: capacity when n_rays_dev is set
  const int* n_rays_dev;    // optional device-side ray count (bounce passes): the launch covers min(*n_rays_dev, N) rays
  // NORMALS
  const uint32_t* wd_bf16;  // density head row as 128 packed bf16x2 words
  float* normals;           // [P,3]
  // BACKWARD
  const float* g_sigma;     // [P]      dL/d sigma (softplus density)
  const float* g_feat;      // [P,16]   dL/d feat (forward feature row layout)
  const float* feat;        // [P,16]   forward outputs
  const float* aux;         // [P,8]    forward aux (mid rgb, raw normal head)
  uint8_t* dy_stash;        // [n_tiles][DY_BLOCKS][16 KB]
  float* g_area;            // [P]      dL/d pixel_area (or sqradius) contribution of each point, or NULL
  int debug;                // RSN_BWD_DEBUG (test build, timing experiments only; results are wrong): 1 = no dY stores,
                            // 8 = no weight streaming
};

struct BBarriers {
  uint64_t w_full[MAX_WSTAGES], w_empty[MAX_WSTAGES];
  uint64_t m_full[NUM_MSTAGES], m_empty[NUM_MSTAGES];
  uint64_t act_ready[8];      // [accumulator buffer of the producing step][64-column group]: two sets, so that the stash warps
                              // may lag the issuer by up to two steps without aliasing the phase parity
  uint64_t a_free[2];         // BACKWARD: the stash warps have read the dY operand that lived in accumulator buffer b
  uint64_t seed_ready;
  uint64_t acc_full[4];       // [accumulator buffer][output half]: wide steps commit columns 0-127 before 128-255
  uint32_t tmem_slot;
};

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void sts128b(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128b(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}
// Stashed ReLU bit masks (csrc/field_layout.cuh): bit i / 16 + i of word w = columns 32 w + 2 i / + 1 of the group.
// 0xFFFF in each half of packed word i whose column was > 0 in the forward.
// (shift bit i & 7 to the top of byte 0 and bit 16 + (i & 7) to the top of byte 2 -- the same shift puts the bits of word
// i + 8 at the top of bytes 1 and 3, so words i and i + 8 share one shifted register --, then one PRMT whose selector
// nibbles 0x8 / 0xA (0x9 / 0xB) replicate the sign of byte 0 / byte 2 (byte 1 / byte 3) over two bytes each)
__device__ __forceinline__ uint32_t relu_mask_word(uint32_t bits, int i) {
  const uint32_t t = bits << (7 - (i & 7));
  uint32_t m;   // (inline PTX: the __byte_perm intrinsic only honours the low three bits of each selector nibble)
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(t), "r"(0u), "r"((i & 8) ? 0xBB99u : 0xAA88u));
  return m;
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// 64 accumulator columns of this row -> (optional ReLU mask from the stashed bit masks) -> bf16 -> the next step's A
// operand: TS form: TMEM columns a_taddr..+31 (over accumulator columns this thread has already read); SS form (test
// build): the shared-memory activation block, in place -- and, when `dy_blk` is given, the row's place in the chunk-major
// dY stash block straight from the registers (in the TS form the stash warps write it, see chain_body).
__device__ __forceinline__ void stg128b(uint8_t* gaddr, uint4 v) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(gaddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <bool MASK, bool TS>
__device__ __forceinline__ void dgrad_group(uint32_t tmem_row_col, uint32_t blk_saddr, uint2 mbits, int row, uint32_t a_taddr,
                                            uint8_t* dy_blk) {
  uint32_t v[2][32];
  tmem_ld32(tmem_row_col, v[0]);
  tmem_ld32(tmem_row_col + 32, v[1]);
  tmem_ld_wait();
  const uint32_t row_saddr = blk_saddr + (uint32_t)row * 128u;
  uint32_t a[32];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t* x = &v[c >> 2][(c & 3) * 8];
    uint4 pk = make_uint4(pack2(__uint_as_float(x[0]), __uint_as_float(x[1])), pack2(__uint_as_float(x[2]), __uint_as_float(x[3])),
                          pack2(__uint_as_float(x[4]), __uint_as_float(x[5])), pack2(__uint_as_float(x[6]), __uint_as_float(x[7])));
    if (MASK) {
      const uint32_t bits = (c >> 2) ? mbits.y : mbits.x;
      const int i0 = (c & 3) * 4;
      pk.x &= relu_mask_word(bits, i0 + 0);
      pk.y &= relu_mask_word(bits, i0 + 1);
      pk.z &= relu_mask_word(bits, i0 + 2);
      pk.w &= relu_mask_word(bits, i0 + 3);
    }
    if (!TS) {
      sts128b(row_saddr + (uint32_t)((c ^ (row & 7)) << 4), pk);
      if (dy_blk) stg128b(dy_blk + stash_chunk_off(row, c), pk);
    } else {
      a[c * 4 + 0] = pk.x, a[c * 4 + 1] = pk.y, a[c * 4 + 2] = pk.z, a[c * 4 + 3] = pk.w;
    }
  }
  if (TS) tmem_st32(a_taddr, a);
}

// Geometry of one frustum sample needed by the area Jacobian: d diag_a / d pixel_area for the contracted
// Gaussian of field.py:90-119 (see frustum_gaussian_contracted in field_fwd.cu; rad_var is linear in
// pixel_area, cov is linear in rad_var, diag = relu(diag(J cov J))).
__device__ __forceinline__ void area_jacobian(const float o[3], const float d[3], float t0, float t1, float (&coef)[3]) {
  const float mu = (t0 + t1) * 0.5f, hw = (t1 - t0) * 0.5f;
  const float hw2 = hw * hw, mu2 = mu * mu, den = 3.0f * mu2 + hw2;
  const float tmean = mu + (2.0f * mu * hw2) / den;
  float m[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) m[a] = o[a] + d[a] * tmean;
  const float c_rad = (mu2 * 0.25f + 0.41666666f * hw2 - 0.26666668f * (hw2 * hw2) / den) * 0.31830987f;  // / pi
  const float dd = fmaxf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2], 1e-10f);
  float nm[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) nm[i][j] = ((i == j) ? 1.0f : 0.0f) - d[i] * (d[j] / dd);
  const float n2 = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
  const float n1 = sqrtf(n2);
  if (n1 > 1.0f) {
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
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) t += J[i][j] * nm[j][k];
        acc += t * J[k][i];
      }
      coef[i] = acc * c_rad;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i) coef[i] = nm[i][i] * c_rad;
  }
}

// Contribution of the 112 d(enc) columns in TMEM (this row) to either
//   NORMALS : d raw / d mean_a = sum_k 2 pi f_k (g_sin E_cos - g_cos E_sin) + g_xyz        (cov constant, Q6)
//   BACKWARD: dL / d diag_a    = sum_k -f_k^2 / 2 (g_sin E_sin + g_cos E_cos)
// where E = the stashed (bf16) encoding values exp(-v/2) sin(s), exp(-v/2) sin(s + pi/2) of the forward.
template <int KIND>
__device__ __forceinline__ void enc_contract(uint32_t tmem_row, uint32_t enc0_saddr, uint32_t enc1_saddr, int row,
                                             float (&out)[3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    uint32_t gs[16], gc[16];
    tmem_ld16(tmem_row + 16 * a, gs);        // d enc of sin features of axis a
    tmem_ld16(tmem_row + 48 + 16 * a, gc);   // d enc of the sin(s + pi/2) features
    // stashed E: sin chunk j = 2a + kk (block 0), cos chunk j = 6 + 2a + kk (block (j >> 3))
    uint4 es[2], ec[2];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int js = 2 * a + kk, jc = 6 + 2 * a + kk;
      es[kk] = lds128b(enc0_saddr + (uint32_t)row * 128u + (uint32_t)(((js & 7) ^ (row & 7)) << 4));
      ec[kk] = lds128b(((jc >> 3) ? enc1_saddr : enc0_saddr) + (uint32_t)row * 128u + (uint32_t)(((jc & 7) ^ (row & 7)) << 4));
    }
    tmem_ld_wait();
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const uint32_t ws = (&es[k >> 3].x)[(k & 7) >> 1], wc = (&ec[k >> 3].x)[(k & 7) >> 1];
      const float Es = (k & 1) ? bf_hi(ws) : bf_lo(ws), Ec = (k & 1) ? bf_hi(wc) : bf_lo(wc);
      const float f = c_freq_b[k];
      if (KIND == KIND_NORMALS)
        acc += f * (__uint_as_float(gs[k]) * Ec - __uint_as_float(gc[k]) * Es);
      else
        acc += (f * f) * (__uint_as_float(gs[k]) * Es + __uint_as_float(gc[k]) * Ec);
    }
    out[a] += (KIND == KIND_NORMALS) ? 6.2831854820251465f * acc : -0.5f * acc;
  }
  if (KIND == KIND_NORMALS) {
    uint32_t gx[16];
    tmem_ld16(tmem_row + 96, gx);
    tmem_ld_wait();
#pragma unroll
    for (int a = 0; a < 3; ++a) out[a] += __uint_as_float(gx[a]);
  }
}

// TS: the dY operand of every step comes from TMEM (tcgen05.mma [d], [a], b-desc), written there by the previous step's
// epilogue; there are no shared-memory activation blocks.  BACKWARD in the TS form launches four more warps (7-10) that read
// each handed-over 64-column group back out of TMEM (it stays valid until the issuer re-uses the accumulator buffer two
// steps later, which waits for a_free) and write it to the dY stash with coalesced st.global.v4 (chunk-major block image,
// field_layout.cuh) -- nothing of the stash sits on the step-critical path.  SS form (test build): the epilogue writes the
// stash rows itself.
template <int KIND, bool TS = false>
__device__ __forceinline__ void chain_body(const BwdParams& p) {
  const int vbid = (int)blockIdx.x, vgrid = (int)gridDim.x;
  constexpr bool STASHW = TS && KIND == KIND_BACKWARD;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ BBarriers bars;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s_act = smem_u32(smem + SM_ACT);
  const uint32_t s_seed = smem_u32(smem + (TS ? SM_SEED_TS : SM_SEED));
  // weight ring: SS form 3 x 32 KB after the activation and seed blocks; TS form 5 x 32 KB from the start of the buffer
  constexpr int NWS = TS ? MAX_WSTAGES : NUM_WSTAGES;
  constexpr int RING_OFF = TS ? SM_ACT : SM_W;
  static_assert(RING_OFF + NWS * W_STAGE_BYTES <= (TS ? SM_SEED_TS : SM_M), "weight ring overlaps what follows it");
  const uint32_t s_w = smem_u32(smem + RING_OFF);
  const uint32_t s_m = smem_u32(smem + SM_M);
  const bool with_enc = (KIND == KIND_NORMALS) || p.want_area;

  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < MAX_WSTAGES; ++i) {
        mbar_init(&bars.w_full[i], 1);
        mbar_init(&bars.w_empty[i], 1);
      }
      for (int i = 0; i < NUM_MSTAGES; ++i) {
        mbar_init(&bars.m_full[i], 1);
        mbar_init(&bars.m_empty[i], TILE);
      }
      for (int i = 0; i < 8; ++i) mbar_init(&bars.act_ready[i], TILE);
      for (int i = 0; i < 2; ++i) mbar_init(&bars.a_free[i], 4);
      mbar_init(&bars.seed_ready, TILE);
      for (int i = 0; i < 4; ++i) mbar_init(&bars.acc_full[i], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&bars.tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_slot;
  const int n_points = (int)rsn_count((int64_t)p.n_points / p.n_samples, p.n_rays_dev) * p.n_samples;
  const int n_tiles = (n_points + TILE - 1) / TILE;
  const int n_my_tiles = (n_tiles > vbid) ? (n_tiles - vbid + vgrid - 1) / vgrid : 0;

  if (warp == 0) {
    // ===================================================================== transposed-weight producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto push = [&](uint32_t off, uint32_t bytes) {
        mbar_wait(&bars.w_empty[stage], phase ^ 1);
        if (p.debug & 8) {
          mbar_arrive(&bars.w_full[stage]);   // timing experiment: no weight traffic
        } else {
          mbar_expect_tx(&bars.w_full[stage], bytes);
          bulk_g2s(smem + RING_OFF + stage * W_STAGE_BYTES, p.wblob_t + off, bytes, &bars.w_full[stage]);
        }
        if (++stage == NWS) {
          stage = 0;
          phase ^= 1;
        }
      };
      for (int it = 0; it < n_my_tiles; ++it) {
        if (KIND == KIND_BACKWARD) {
          push(BT_RGB, 16384);
          for (int c = 0; c < 2; ++c) push(BT_MID + c * 32768, 32768);
          for (int c = 0; c < 4; ++c) push(BT_BOTT + c * 32768, 32768);
          push(BT_HEADS, 32768);
        }
        for (int l = 7; l >= 1; --l) {
          for (int c = 0; c < 4; ++c) push(BT_L(l) + c * 32768, 32768);
          if (l == 4 && with_enc)
            for (int c = 0; c < 4; ++c) push(BT_L4E + c * 16384, 16384);
        }
        if (with_enc)
          for (int c = 0; c < 4; ++c) push(BT_L0 + c * 16384, 16384);
      }
    }
  } else if (warp == 6) {
    // ===================================================================== stashed-encoding producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_my_tiles; ++it) {
        const int tile = vbid + it * vgrid;
        const uint8_t* xt = p.x_stash + (size_t)tile * STASH_TILE_BYTES;
        auto push = [&](int blk) {
          mbar_wait(&bars.m_empty[stage], phase ^ 1);
          mbar_expect_tx(&bars.m_full[stage], BLOCK_BYTES);
          bulk_g2s(smem + SM_M + stage * BLOCK_BYTES, xt + (size_t)blk * BLOCK_BYTES, BLOCK_BYTES, &bars.m_full[stage]);
          if (++stage == NUM_MSTAGES) {
            stage = 0;
            phase ^= 1;
          }
        };
        if (with_enc) {       // before the layer-4 epilogue and again before the layer-0 epilogue
          push(STASH_ENC);
          push(STASH_ENC + 1);
          push(STASH_ENC);
          push(STASH_ENC + 1);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // (whole warp: control flow and barrier waits; tcgen05.mma / tcgen05.commit under elect_one_sync(), see umma.cuh)
    {
      int stage = 0;
      uint32_t wphase = 0, ar_phase = 0;
      auto mma_commit = [&](uint64_t* bar) {
        if (elect_one_sync()) umma::mma_commit(bar);
      };
      int buf = 0;
      constexpr uint32_t ID256 = instr_desc_bf16(128, 256, 0, 0);
      constexpr uint32_t ID128 = instr_desc_bf16(128, 128, 0, 0);
      auto ring_wait = [&]() -> uint32_t {
        mbar_wait(&bars.w_full[stage], wphase);
        tc_fence_after();
        return s_w + (uint32_t)stage * W_STAGE_BYTES;
      };
      auto ring_advance = [&]() -> int {     // -> the slot just consumed (to be released with ring_release_slot)
        const int s_ = stage;
        if (++stage == NWS) {
          stage = 0;
          wphase ^= 1;
        }
        return s_;
      };
      auto ring_release_slot = [&](int s_) { mma_commit(&bars.w_empty[s_]); };
      auto ring_release = [&]() { ring_release_slot(ring_advance()); };
      // one commit for both output halves (steps that are not split)
      auto commit_acc = [&]() {
        mma_commit(&bars.acc_full[2 * buf]);
        mma_commit(&bars.acc_full[2 * buf + 1]);
      };
      auto wait_act = [&](int g) {
        const int i = (buf ^ 1) * 4 + g;     // handed over by the previous step, which accumulated into the other buffer
        mbar_wait(&bars.act_ready[i], (ar_phase >> i) & 1u);
        ar_phase ^= (1u << i);
        tc_fence_after();
      };
      int use = 0;     // accumulator-buffer uses so far (use u accumulates into buffer u & 1)
      // STASHW: before the first MMA of use u overwrites buffer u & 1, the stash warps must have read the dY operand that
      // the epilogue of use u - 2 left in its first 128 columns
      auto wait_a_free = [&]() {
        if (STASHW && use >= 2) mbar_wait(&bars.a_free[use & 1], (uint32_t)((use - 2) >> 1) & 1u);
        ++use;
      };
      constexpr uint32_t HI = desc_hi_sw128(1024);
      auto issue_kb = [&](uint32_t a_addr, uint32_t b_addr, int ksteps, uint32_t idesc, uint32_t tmem_d, bool& acc) {
        const uint32_t a_lo = desc_lo(a_addr, 16), b_lo = desc_lo(b_addr, 16);
        if (elect_one_sync()) {
          mma_bf16_ss_lo(tmem_d, a_lo, b_lo, HI, idesc, acc ? 1u : 0u);
          if (ksteps == 4) {
            mma_bf16_ss_lo(tmem_d, a_lo + 2, b_lo + 2, HI, idesc, 1u);
            mma_bf16_ss_lo(tmem_d, a_lo + 4, b_lo + 4, HI, idesc, 1u);
            mma_bf16_ss_lo(tmem_d, a_lo + 6, b_lo + 6, HI, idesc, 1u);
          }
        }
        acc = true;
      };
      // K-block g of the current dY: shared-memory block g (SS) or TMEM columns [32 g, 32 g + 32) of the OTHER
      // accumulator buffer, where the previous step's epilogue left it (TS)
      auto issue_act = [&](int g, uint32_t b_addr, uint32_t idesc, uint32_t tmem_d, bool& acc) {
        if (!TS) {
          issue_kb(s_act + g * BLOCK_BYTES, b_addr, 4, idesc, tmem_d, acc);
        } else {
          const uint32_t b_lo = desc_lo(b_addr, 16), a0 = tmem + (uint32_t)(buf ^ 1) * 256 + (uint32_t)g * 32u;
          if (elect_one_sync()) {
            mma_bf16_ts_lo(tmem_d, a0, b_lo, HI, idesc, acc ? 1u : 0u);
            mma_bf16_ts_lo(tmem_d, a0 + 8, b_lo + 2, HI, idesc, 1u);
            mma_bf16_ts_lo(tmem_d, a0 + 16, b_lo + 4, HI, idesc, 1u);
            mma_bf16_ts_lo(tmem_d, a0 + 24, b_lo + 6, HI, idesc, 1u);
          }
          acc = true;
        }
      };
      constexpr uint32_t HALF_B = 128u * 128u;   // bytes of 128 weight rows (one output half) inside a [256][64] K-block image
      // NORMALS only (measured, 30 launches alternating between two builds: normals 1.97 -> 1.88 ms, but BACKWARD 2.98 ->
      // 3.09 ms, + area 3.50 -> 3.64 ms: there the stash warps' TMEM reads and stores already pace the epilogue, and the
      // early first half only adds issue work).
      constexpr bool SPLIT = KIND == KIND_NORMALS || RSN_BWD_SPLIT;
      // A wide step (K = 256 from the four handed-over groups, N = 256) as two N = 128 accumulations (columns 0-127 | 128-255,
      // weight rows 0-127 | 128-255 of every K-block image), the first COMMITTED while the tensor pipe still works on the
      // second -- the epilogue converts groups 0, 1 (and the next step's first K-blocks start) two K-block times earlier;
      // same tensor time as N = 256 (the forward kernel's form, field_fwd.cu).  Issue order:
      //   [kb0: h0 h1] [kb1: h0 h1] [kb2: h0] [kb3: h0] before_h0() commit(h0) [kb2: h1] [kb3: h1] before_h1() commit(h1)
      // before_h0 / before_h1: extra MMAs of the step that belong in front of the first / second commit.
      auto issue_wide = [&](auto&& before_h0, auto&& before_h1) {
        const uint32_t tm = tmem + (uint32_t)buf * 256;
        bool acc0 = false, acc1 = false;
        for (int g = 0; g < 2; ++g) {
          wait_act(g);
          const uint32_t w = ring_wait();
          issue_act(g, w, ID128, tm, acc0);
          issue_act(g, w + HALF_B, ID128, tm + 128, acc1);
          ring_release();
        }
        wait_act(2);
        const uint32_t w2 = ring_wait();
        const int s2 = ring_advance();
        issue_act(2, w2, ID128, tm, acc0);
        wait_act(3);
        const uint32_t w3 = ring_wait();
        const int s3 = ring_advance();
        issue_act(3, w3, ID128, tm, acc0);
        before_h0(acc0);
        mma_commit(&bars.acc_full[2 * buf]);
        issue_act(2, w2 + HALF_B, ID128, tm + 128, acc1);
        ring_release_slot(s2);
        issue_act(3, w3 + HALF_B, ID128, tm + 128, acc1);
        ring_release_slot(s3);
        before_h1(acc1);
        mma_commit(&bars.acc_full[2 * buf + 1]);
      };
      auto nothing = [](bool&) {};
      constexpr uint32_t ENC4_COL = TS ? 128u : 0u;   // layer-4 encoding part: columns of the other buffer (TS: A sits in 0..127)
      for (int it = 0; it < n_my_tiles; ++it) {
        bool acc;
        if (KIND == KIND_BACKWARD) {
          // S0: d mid_hidden = dY_rgb (K=16) x Wrgb^T, N = 128
          mbar_wait(&bars.seed_ready, (uint32_t)it & 1u);
          tc_fence_after();
          acc = false;
          wait_a_free();
          uint32_t w = ring_wait();
          issue_kb(s_seed, w, 1, ID128, tmem + (uint32_t)buf * 256, acc);
          ring_release();
          commit_acc();
          buf ^= 1;
          // S1: d bottleneck = dY_mid (K=128) x Wmid[:,34:]^T, N = 256
          acc = false;
          wait_a_free();
          for (int g = 0; g < 2; ++g) {
            wait_act(g);
            w = ring_wait();
            issue_act(g, w, ID256, tmem + (uint32_t)buf * 256, acc);
            ring_release();
          }
          commit_acc();
          buf ^= 1;
          // S2: d emb = dY_bott (K=256) x Wbott^T + dY_heads (K=16, seed block columns 16-31) x Wheads^T
          acc = false;
          wait_a_free();
          for (int g = 0; g < 4; ++g) {
            wait_act(g);
            w = ring_wait();
            issue_act(g, w, ID256, tmem + (uint32_t)buf * 256, acc);
            ring_release();
          }
          w = ring_wait();
          issue_kb(s_seed + 32, w, 1, ID256, tmem + (uint32_t)buf * 256, acc);
          ring_release();
          commit_acc();
          buf ^= 1;
        }
        // base layers 7..1: d h_{l-1} = dY_l x W_l^T  (layer 4: hidden part, then the encoding part)
        for (int l = 7; l >= 1; --l) {
          wait_a_free();
          if (!SPLIT || (l == 4 && with_enc)) {
            // Not split (layer 4 with its encoding part): the encoding part of layer 4 (four more weight chunks, N = 128 into the OTHER buffer) has to be
            // complete before the epilogue hands group 0 over (the next step then starts to overwrite that buffer), and
            // the round-robin weight ring cannot deliver its fourth chunk while two slots of the wide step are still held.
            // (every act_ready of the previous epilogue has been observed => it no longer reads buf^1)
            bool acc = false;
            for (int g = 0; g < 4; ++g) {
              wait_act(g);
              const uint32_t w = ring_wait();
              issue_act(g, w, ID256, tmem + (uint32_t)buf * 256, acc);
              ring_release();
            }
            if (l == 4 && with_enc) {
              bool acc_e = false;
              for (int g = 0; g < 4; ++g) {
                const uint32_t w = ring_wait();
                issue_act(g, w, ID128, tmem + (uint32_t)(buf ^ 1) * 256 + ENC4_COL, acc_e);
                ring_release();
              }
            }
            commit_acc();
          } else {
            issue_wide(nothing, nothing);
          }
          buf ^= 1;
        }
        if (!with_enc) {
          // nobody consumes dY_0 as an operand: still observe its act_ready completions so that the parity
          // bookkeeping of the next tile stays in step with the barriers
          for (int g = 0; g < 4; ++g) wait_act(g);
        }
        if (with_enc) {
          // layer 0: d enc = dY_0 x W0^T, N = 128
          acc = false;
          wait_a_free();
          for (int g = 0; g < 4; ++g) {
            wait_act(g);
            const uint32_t w = ring_wait();
            issue_act(g, w, ID128, tmem + (uint32_t)buf * 256, acc);
            ring_release();
          }
          commit_acc();
          buf ^= 1;
        }
      }
    }
  } else if (warp >= 2 && warp < 6) {
    // ===================================================================== epilogue warps (thread = point row)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t af_phase = 0, mphase = 0;
    int buf = 0, mstage = 0;
    for (int it = 0; it < n_my_tiles; ++it) {
      const int tile = vbid + it * vgrid;
      const int pt = tile * TILE + row;
      const bool valid = pt < n_points;
      uint8_t* const dyt = (KIND == KIND_BACKWARD) ? p.dy_stash + (size_t)tile * DY_BLOCKS * BLOCK_BYTES : nullptr;
      auto dblk = [&](int b) -> uint8_t* { return dyt + (size_t)b * BLOCK_BYTES; };
      auto wait_half = [&](int h) {          // output half h of the step accumulating into `buf` is complete
        const int i = 2 * buf + h;
        mbar_wait(&bars.acc_full[i], (af_phase >> i) & 1u);
        af_phase ^= (1u << i);
        tc_fence_after();
      };
      auto wait_acc = [&]() {                // the whole accumulator (steps that are not split commit both halves at once)
        wait_half(0);
        wait_half(1);
      };
      // one 64-column group of a step: convert and hand to the issuer (and, BACKWARD, to the stash warps): act_ready[buf][g]
      auto convert = [&](auto mask_c, int g, uint2 mbits, uint8_t* stash_blk) {
        constexpr bool MASK = decltype(mask_c)::value;
        const uint32_t acc_c = tlane + (uint32_t)buf * 256 + g * 64, a_t = tlane + (uint32_t)buf * 256 + g * 32;
        dgrad_group<MASK, TS>(acc_c, s_act + g * BLOCK_BYTES, mbits, row, a_t, (p.debug & 1) ? nullptr : stash_blk);
        if (TS) tmem_st_wait(); else fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&bars.act_ready[buf * 4 + g]);
      };
      using T_ = std::integral_constant<bool, true>;
      using F_ = std::integral_constant<bool, false>;
      auto mask_wait = [&]() -> uint32_t {
        mbar_wait(&bars.m_full[mstage], mphase);
        return s_m + (uint32_t)mstage * BLOCK_BYTES;
      };
      auto mask_release = [&]() {
        mbar_arrive(&bars.m_empty[mstage]);
        if (++mstage == NUM_MSTAGES) {
          mstage = 0;
          mphase ^= 1;
        }
      };
      float red[3] = {0.f, 0.f, 0.f};   // d mean (NORMALS) or dL/d diag (BACKWARD), summed over both enc GEMMs
      // ReLU bit masks of this row (8 bytes per layer and group), fetched before the accumulator wait of the step
      const uint2* const mrow = reinterpret_cast<const uint2*>(p.x_stash + (size_t)tile * STASH_TILE_BYTES + STASH_MASK_OFF);
      auto load_masks = [&](int layer, int ngroups, uint2 (&mk)[4]) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (g < ngroups) mk[g] = __ldg(mrow + mask_entry(layer, g, row));
      };

      if (KIND == KIND_BACKWARD) {
        // ---- seed: head pre-activation gradients of this point -> seed block (and the dY stash)
        float dr[3] = {0.f, 0.f, 0.f}, dh[11];
#pragma unroll
        for (int i = 0; i < 11; ++i) dh[i] = 0.f;
        if (valid) {
          const float4* f4 = reinterpret_cast<const float4*>(p.feat + (size_t)pt * N_FEAT);
          const float4* g4 = reinterpret_cast<const float4*>(p.g_feat + (size_t)pt * N_FEAT);
          const float4 f0 = __ldg(f4), f1 = __ldg(f4 + 1), f2 = __ldg(f4 + 2), f3 = __ldg(f4 + 3);
          const float4 g0 = __ldg(g4), g1 = __ldg(g4 + 1), g2 = __ldg(g4 + 2), g3 = __ldg(g4 + 3);
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.aux + (size_t)pt * 8));
          const float4 a1 = __ldg(reinterpret_cast<const float4*>(p.aux + (size_t)pt * 8) + 1);
          const float mid[3] = {a0.x, a0.y, a0.z}, rawn[3] = {a0.w, a1.x, a1.y};
          const float diff[3] = {f0.w, f1.x, f1.y}, tint[3] = {f1.z, f1.w, f2.x}, pn[3] = {f2.y, f2.z, f2.w};
          const float g_rgb[3] = {g0.x, g0.y, g0.z}, g_diff[3] = {g0.w, g1.x, g1.y}, g_tint[3] = {g1.z, g1.w, g2.x};
          float g_pn[3] = {g2.y, g2.z, g2.w};
          const float rs = f3.x, raw = f3.z, g_rs = g3.x, g_ndd = g3.y;
          if (p.mode == 0) {
            const int ray = pt / p.n_samples;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              g_pn[a] += g_ndd * __ldg(p.dirs + (size_t)ray * 3 + a);   // n.d = sum d_a pn_a (field.py:204)
              dr[a] = g_rgb[a] * tint[a] * mid[a] * (1.f - mid[a]);
              dh[5 + a] = (g_rgb[a] + g_diff[a]) * diff[a] * (1.f - diff[a]);
              dh[8 + a] = (g_rgb[a] * mid[a] + g_tint[a]) * tint[a] * (1.f - tint[a]);
            }
            const float z = raw + 0.5f;
            dh[0] = (p.g_sigma ? __ldg(p.g_sigma + pt) : 0.f) * (z > 20.f ? 1.f : 1.f / (1.f + __expf(-z)));   // softplus'
            // pn = normalize(-normalize(v)):  d v = -(I - pn pn^T) d pn / |v|
            const float nv = fmaxf(sqrtf(rawn[0] * rawn[0] + rawn[1] * rawn[1] + rawn[2] * rawn[2]), 1e-12f);
            const float dot = pn[0] * g_pn[0] + pn[1] * g_pn[1] + pn[2] * g_pn[2];
#pragma unroll
            for (int a = 0; a < 3; ++a) dh[1 + a] = -(g_pn[a] - pn[a] * dot) / nv;
            dh[4] = g_rs * rs * (1.f - rs);
          } else {
#pragma unroll
            for (int a = 0; a < 3; ++a) dr[a] = g_rgb[a] * mid[a] * (1.f - mid[a]);   // infinity colour = mid
          }
        }
        const uint4 c0 = make_uint4(pack2(dr[0], dr[1]), pack2(dr[2], 0.f), 0u, 0u);
        const uint4 c2 = make_uint4(pack2(dh[0], dh[1]), pack2(dh[2], dh[3]), pack2(dh[4], dh[5]), pack2(dh[6], dh[7]));
        const uint4 c3 = make_uint4(pack2(dh[8], dh[9]), pack2(dh[10], 0.f), 0u, 0u);
        const uint4 zz = make_uint4(0u, 0u, 0u, 0u);
        // (the previous tile's S0 / S2 have read the block: its last accumulator was waited for)
        const uint32_t srow = s_seed + (uint32_t)row * 128u;
        uint8_t* const sdst = dblk(DY_SEED);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 v = (c == 0) ? c0 : (c == 2) ? c2 : (c == 3) ? c3 : zz;
          sts128b(srow + (uint32_t)((c ^ (row & 7)) << 4), v);            // swizzled: the A operand of S0 / S2
          if (!(p.debug & 1)) stg128b(sdst + stash_chunk_off(row, c), v);  // chunk-major: the dY stash
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&bars.seed_ready);
        // ---- E0: d mid_hidden * (mid_hidden > 0) -> dY_mid (act blocks 0,1)
        uint2 mk0[4];
        load_masks(8, 2, mk0);
        wait_acc();
        for (int g = 0; g < 2; ++g) convert(T_{}, g, mk0[g], dblk(DY_MID + g));
        buf ^= 1;
        // ---- E1: d bottleneck (no activation) -> dY_bott
        wait_acc();
        // (dY_bott is not stashed: the wgrad derives the bottleneck layer's gradients from dY_mid and h7)
        for (int g = 0; g < 4; ++g) convert(F_{}, g, make_uint2(0u, 0u), nullptr);
        buf ^= 1;
      } else {
        // ---- NORMALS seed: dY_7 = w_density * (h7 > 0)
        uint2 mk7[4];
        load_masks(7, 4, mk7);
        for (int g = 0; g < 4; ++g) {
          const uint32_t row_saddr = s_act + g * BLOCK_BYTES + (uint32_t)row * 128u;
          uint32_t a[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 wv = __ldg(reinterpret_cast<const uint4*>(p.wd_bf16) + g * 8 + c);
            const uint32_t bits = (c >> 2) ? mk7[g].y : mk7[g].x;
            const int i0 = (c & 3) * 4;
            const uint4 m = make_uint4(wv.x & relu_mask_word(bits, i0), wv.y & relu_mask_word(bits, i0 + 1),
                                       wv.z & relu_mask_word(bits, i0 + 2), wv.w & relu_mask_word(bits, i0 + 3));
            if (!TS) sts128b(row_saddr + (uint32_t)((c ^ (row & 7)) << 4), m);
            else a[c * 4] = m.x, a[c * 4 + 1] = m.y, a[c * 4 + 2] = m.z, a[c * 4 + 3] = m.w;
          }
          if (TS) {   // the first step accumulates into `buf`: its A operand goes to the other buffer
            tmem_st32(tlane + (uint32_t)(buf ^ 1) * 256 + g * 32, a);
            tmem_st_wait();
          } else {
            fence_proxy_async();
          }
          tc_fence_before();
          mbar_arrive(&bars.act_ready[(buf ^ 1) * 4 + g]);
        }
      }
      // ---- chain: (BACKWARD: E2 = d emb) then layers 7..1; each: acc * (h_{l-1} > 0) -> dY_{l-1}
      const int first = (KIND == KIND_BACKWARD) ? 8 : 7;
      for (int l = first; l >= 1; --l) {
        uint2 mk[4];
        load_masks(l - 1, 4, mk);       // this step masks with h_{l-1}
        wait_half(0);                   // wide step: columns 0-127 first (groups 0, 1), see issue_wide
        if (l == 4 && with_enc) {
          // encoding part of layer 4 (other accumulator, columns 0..111): consume before the hidden part
          const uint32_t e0 = mask_wait();
          // both enc blocks are needed at once: they occupy two consecutive ring stages
          uint32_t mphase1 = mphase;
          int mstage1 = mstage + 1;
          if (mstage1 == NUM_MSTAGES) {
            mstage1 = 0;
            mphase1 ^= 1;
          }
          mbar_wait(&bars.m_full[mstage1], mphase1);
          const uint32_t e1 = s_m + (uint32_t)mstage1 * BLOCK_BYTES;
          enc_contract<KIND>(tlane + (uint32_t)(buf ^ 1) * 256 + (TS ? 128u : 0u), e0, e1, row, red);
          mask_release();
          mask_release();
        }
        for (int g = 0; g < 4; ++g) { // output = dY_{l-1}: stash block DY_H + 4 (l-1) + g
          if (g == 2) wait_half(1);
          convert(T_{}, g, mk[g], (KIND == KIND_BACKWARD) ? dblk(DY_H + 4 * (l - 1) + g) : nullptr);
        }
        buf ^= 1;
      }
      if (with_enc) {
        wait_acc();
        const uint32_t e0 = mask_wait();
        uint32_t mphase1 = mphase;
        int mstage1 = mstage + 1;
        if (mstage1 == NUM_MSTAGES) {
          mstage1 = 0;
          mphase1 ^= 1;
        }
        mbar_wait(&bars.m_full[mstage1], mphase1);
        const uint32_t e1 = s_m + (uint32_t)mstage1 * BLOCK_BYTES;
        enc_contract<KIND>(tlane + (uint32_t)buf * 256, e0, e1, row, red);
        mask_release();
        mask_release();
        tc_fence_before();
        buf ^= 1;
        if (valid) {
          if (KIND == KIND_NORMALS) {
            // Field.get_normals: -F.normalize(grad)   (F.normalize: x / max(|x|, 1e-12))
            const float nn = fmaxf(sqrtf(red[0] * red[0] + red[1] * red[1] + red[2] * red[2]), 1e-12f);
            p.normals[(size_t)pt * 3 + 0] = -(red[0] / nn);
            p.normals[(size_t)pt * 3 + 1] = -(red[1] / nn);
            p.normals[(size_t)pt * 3 + 2] = -(red[2] / nn);
          } else if (p.g_area) {
            float coef[3];
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
              area_jacobian(o, d, __ldg(b), __ldg(b + 1), coef);
            } else {
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                const float w = __ldg(p.dirs + (size_t)pt * 3 + a);
                coef[a] = 0.6f * (1.0f - w * w);   // cov = 0.6 sqradius (I - w w^T)  (field.py:196)
              }
            }
            // relu on the diagonal (field.py:113-115): a clamped entry passes no gradient.  coef >= 0 and
            // pixel_area > 0 make the unclamped diagonal positive except for rounding; treat it as open.
            p.g_area[pt] = red[0] * coef[0] + red[1] * coef[1] + red[2] * coef[2];
          }
        }
      }
    }
  }
  if constexpr (STASHW) {
    if (warp >= 7) {
      // ===================================================================== dY stash warps
      const int q = warp & 3;                                   // TMEM lane quarter this warp may touch
      const int row = q * 32 + lane;
      const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
      uint32_t ar_phase = 0;                                    // bit 4 b + g: parity of the next completion of act_ready[4 b + g]
      int sbuf = 0;                                             // accumulator buffer of the step being stashed
      for (int it = 0; it < n_my_tiles; ++it) {
        uint8_t* const dyt = p.dy_stash + (size_t)(vbid + it * vgrid) * DY_BLOCKS * BLOCK_BYTES;
        // one step: `ng` groups handed over by its epilogue; stashed to blocks blk0 .. (blk0 < 0: only observed)
        // RSN_STASH_LAG_BWD = k: group g is read and stored once the epilogue has handed group min(g + k, last) over (4: the
        // whole step after its last hand-over).  The reads and stores then fall into the stretch in which the epilogue
        // warps (same SMSPs) wait for the next step's accumulator instead of competing with their conversion of the
        // following group: backward 2.99 -> 2.84 ms, + area 3.49 -> 3.39 ms (30 launches alternating between two builds).
        // The operand stays valid until the issuer re-uses the buffer two steps later (a_free).
        auto follow = [&](int ng, int blk0) {
          int waited = 0;
          for (int g = 0; g < ng; ++g) {
            const int need = min(g + RSN_STASH_LAG_BWD, ng - 1);
            while (waited <= need) {
              const int i = sbuf * 4 + waited;
              mbar_wait(&bars.act_ready[i], (ar_phase >> i) & 1u);
              ar_phase ^= (1u << i);
              ++waited;
            }
            if (blk0 < 0 || (p.debug & 1)) continue;
            tc_fence_after();
            uint32_t a[32];
            tmem_ld32(tlane + (uint32_t)sbuf * 256 + (uint32_t)g * 32u, a);
            tmem_ld_wait();
            uint8_t* const dst = dyt + (size_t)(blk0 + g) * BLOCK_BYTES + stash_chunk_off(row, 0);
#pragma unroll
            for (int c = 0; c < 8; ++c) stg128b(dst + c * STASH_CHUNK_STRIDE, make_uint4(a[c * 4], a[c * 4 + 1], a[c * 4 + 2], a[c * 4 + 3]));
          }
          tc_fence_before();                                    // this warp no longer reads the operand in buffer sbuf
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.a_free[sbuf]);
          sbuf ^= 1;
        };
        follow(2, DY_MID);                                      // S0 -> E0: dY_mid
        follow(4, -1);                                          // S1 -> E1: d bottleneck, not stashed
        for (int l = 8; l >= 1; --l) follow(4, DY_H + 4 * (l - 1));   // S2, layers 7..1 -> dY_{l-1}
        if (with_enc) follow(0, -1);                            // layer 0 (encoding Jacobian): hands nothing over
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int KIND, bool TS>
__global__ void __launch_bounds__((TS && KIND == KIND_BACKWARD) ? B_THREADS_STASH : B_THREADS, 1)
    field_chain_kernel(const __grid_constant__ BwdParams p) {
  chain_body<KIND, TS>(p);
}

template <int KIND>
int launch_chain(const BwdParams& p, cudaStream_t stream) {
  const size_t smem = SM_TOTAL + 1024;
  static std::atomic<unsigned long long> done[2];
  RSN_CUDA(rsn_ensure_smem(field_chain_kernel<KIND, true>, (int)smem, done[0]));
  const int grid = std::min(p.n_tiles, rsn_num_sms());
#ifdef RSN_DEBUG_SWITCHES
  // RSN_BWD_TS=0 selects the shared-memory (SS) operand form (same results bit for bit; test build only)
  if (rsn_env_int("RSN_BWD_TS", 1) == 0) {
    RSN_CUDA(rsn_ensure_smem(field_chain_kernel<KIND, false>, (int)smem, done[1]));
    field_chain_kernel<KIND, false><<<grid, B_THREADS, smem, stream>>>(p);
    RSN_LAUNCH_CHECK("field_chain_kernel");
    return 0;
  }
#endif
  field_chain_kernel<KIND, true><<<grid, KIND == KIND_BACKWARD ? B_THREADS_STASH : B_THREADS, smem, stream>>>(p);
  RSN_LAUNCH_CHECK("field_chain_kernel");
  return 0;
}

}  // namespace
