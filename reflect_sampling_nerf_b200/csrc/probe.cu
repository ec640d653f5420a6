// Unit probes for the tcgen05 building blocks in umma.cuh.  They are exported through the C-ABI and
// exercised by tests/test_umma_probe.py so that a descriptor / swizzle / TMEM-mapping mistake shows up in
// a 128-row single-tile GEMM rather than inside the fused field kernels.
//
//   rsn_probe_umma_kmajor : D[128,N]  = X[128,K] * W[N,K]^T     (forward / dgrad operand form)
//   rsn_probe_umma_mnmajor: D[M,N]    = U[128,M]^T * V[128,N]   (wgrad operand form; K = 128 points)
// Operands arrive as pre-swizzled block images (umma::block_off) exactly as the field kernels stage them.
#include "rsn_common.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

__global__ void __launch_bounds__(160, 1) probe_kmajor_kernel(const uint8_t* __restrict__ x_blocks,
                                                              const uint8_t* __restrict__ w_blocks, int N, int KB,
                                                              int n_split, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_slot;
  uint8_t* sx = smem;
  uint8_t* sw = smem + (size_t)KB * 16384;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w_block_bytes = (uint32_t)N * 128u;

  if (warp == 4) {
    if (lane == 0) {
      mbar_init(&bar_load, 1);
      mbar_init(&bar_mma, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_base_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;

  if (warp == 4 && lane == 0) {
    mbar_expect_tx(&bar_load, (uint32_t)KB * (16384u + w_block_bytes));
    for (int kb = 0; kb < KB; ++kb) {
      bulk_g2s(sx + (size_t)kb * 16384, x_blocks + (size_t)kb * 16384, 16384, &bar_load);
      bulk_g2s(sw + (size_t)kb * w_block_bytes, w_blocks + (size_t)kb * w_block_bytes, w_block_bytes, &bar_load);
    }
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const int n_inst = N / n_split;
    const uint32_t idesc = instr_desc_bf16(128, n_inst, 0, 0);
    for (int h = 0; h < n_split; ++h) {
      for (int kb = 0; kb < KB; ++kb) {
        for (int k = 0; k < 4; ++k) {
          uint64_t da = smem_desc_sw128(smem_u32(sx + (size_t)kb * 16384) + k * 32, 16, 1024);
          uint64_t db = smem_desc_sw128(smem_u32(sw + (size_t)kb * w_block_bytes) + h * n_inst * 128 + k * 32, 16, 1024);
          mma_bf16_ss(tmem + h * n_inst, da, db, idesc, (kb | k) != 0);
        }
      }
    }
    mma_commit(&bar_mma);
  }
  if (warp < 4) {
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) out[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

__global__ void __launch_bounds__(160, 1) probe_mnmajor_kernel(const uint8_t* __restrict__ u_blocks,
                                                               const uint8_t* __restrict__ v_blocks, int MB, int NB,
                                                               float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_slot;
  uint8_t* su = smem;
  uint8_t* sv = smem + (size_t)MB * 16384;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = MB * 64, N = NB * 64;

  if (warp == 4) {
    if (lane == 0) {
      mbar_init(&bar_load, 1);
      mbar_init(&bar_mma, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_base_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;

  if (warp == 4 && lane == 0) {
    mbar_expect_tx(&bar_load, (uint32_t)(MB + NB) * 16384u);
    for (int b = 0; b < MB; ++b) bulk_g2s(su + (size_t)b * 16384, u_blocks + (size_t)b * 16384, 16384, &bar_load);
    for (int b = 0; b < NB; ++b) bulk_g2s(sv + (size_t)b * 16384, v_blocks + (size_t)b * 16384, 16384, &bar_load);
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = instr_desc_bf16(M, N, 1, 1);
    for (int k = 0; k < 8; ++k) {  // 16 points (= 2 swizzle atoms of 8 rows) per instruction
      uint64_t da = smem_desc_sw128(smem_u32(su) + k * 2048, 16384, 1024);
      uint64_t db = smem_desc_sw128(smem_u32(sv) + k * 2048, 16384, 1024);
      mma_bf16_ss(tmem, da, db, idesc, k != 0);
    }
    mma_commit(&bar_mma);
  }
  if (warp < 4) {
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      if (row < M) {
#pragma unroll
        for (int j = 0; j < 16; ++j) out[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

}  // namespace

extern "C" int rsn_probe_umma_kmajor(const void* x_blocks, const void* w_blocks, int64_t n_out, int64_t k_blocks,
                                     int64_t n_split, float* out, cudaStream_t stream) {
  RSN_ARG(n_out >= 16 && n_out <= 256 && n_out % 16 == 0, "rsn_probe_umma_kmajor: n_out in [16,256], multiple of 16");
  RSN_ARG(k_blocks >= 1 && k_blocks <= 4, "rsn_probe_umma_kmajor: k_blocks in [1,4]");
  RSN_ARG(n_split == 1 || (n_split == 2 && n_out % 32 == 0), "rsn_probe_umma_kmajor: n_split 1 or 2");
  size_t smem = (size_t)k_blocks * (16384 + n_out * 128) + 1024;
  RSN_CUDA(cudaFuncSetAttribute(probe_kmajor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kmajor_kernel<<<1, 160, smem, stream>>>((const uint8_t*)x_blocks, (const uint8_t*)w_blocks, (int)n_out,
                                                (int)k_blocks, (int)n_split, out);
  RSN_LAUNCH_CHECK("probe_kmajor_kernel");
  return 0;
}

extern "C" int rsn_probe_umma_mnmajor(const void* u_blocks, const void* v_blocks, int64_t m_blocks, int64_t n_blocks,
                                      float* out, cudaStream_t stream) {
  RSN_ARG(m_blocks == 2, "rsn_probe_umma_mnmajor: m_blocks must be 2 (M = 128)");
  RSN_ARG(n_blocks >= 1 && n_blocks <= 4, "rsn_probe_umma_mnmajor: n_blocks in [1,4]");
  size_t smem = (size_t)(m_blocks + n_blocks) * 16384 + 1024;
  RSN_CUDA(cudaFuncSetAttribute(probe_mnmajor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_mnmajor_kernel<<<1, 160, smem, stream>>>((const uint8_t*)u_blocks, (const uint8_t*)v_blocks, (int)m_blocks,
                                                 (int)n_blocks, out);
  RSN_LAUNCH_CHECK("probe_mnmajor_kernel");
  return 0;
}

// ---- the wgrad's operand form for chunk-major stash blocks (field_layout.cuh): NO-swizzle MN-major descriptors.
// D[M = 128, N = 64 NB] = U[128 points, 128]^T * V[128 points, 64 NB]; U / V arrive as chunk-major block images (two
// 64-point slabs per block).  Per slab the blocks sit in shared memory the way the wgrad stages them: [MB + NB blocks][8 KB],
// so the 16-byte chunk column j (8 features) of the slab's A (or B) operand starts at j * 1024 and its points follow at
// 16 B: core matrix (8 points x 8 features) = 128 contiguous bytes, `mn_stride` = 1024 between core matrices along the
// features, `k_stride` = 128 along the points; one K = 16 step advances the start address by 256 B.
// `lbo` / `sbo` are passed in so that the test pins which descriptor field carries which stride.
namespace {
__global__ void __launch_bounds__(160, 1) probe_mnmajor_cm_kernel(const uint8_t* __restrict__ u_blocks,
                                                                  const uint8_t* __restrict__ v_blocks, int NB, int lbo, int sbo,
                                                                  float* __restrict__ out, int iters, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int MB = 2;
  const int N = NB * 64;
  const int slab_bytes = (MB + NB) * 8192;
  if (warp == 4) {
    if (lane == 0) {
      mbar_init(&bar_load, 1);
      mbar_init(&bar_mma, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_base_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  if (warp == 4 && lane == 0) {
    mbar_expect_tx(&bar_load, 2u * (uint32_t)slab_bytes);
    for (int s = 0; s < 2; ++s) {
      for (int b = 0; b < MB; ++b) bulk_g2s(smem + s * slab_bytes + b * 8192, u_blocks + (size_t)b * 16384 + s * 8192, 8192, &bar_load);
      for (int b = 0; b < NB; ++b)
        bulk_g2s(smem + s * slab_bytes + (MB + b) * 8192, v_blocks + (size_t)b * 16384 + s * 8192, 8192, &bar_load);
    }
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = instr_desc_bf16(128, N, 1, 1);
    const long long t0 = clock64();
    for (int it = 0; it < (iters > 0 ? iters : 1); ++it)
      for (int s = 0; s < 2; ++s)
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = smem_desc_noswz(smem_u32(smem + s * slab_bytes) + k * 256, lbo, sbo);
          const uint64_t db = smem_desc_noswz(smem_u32(smem + s * slab_bytes + MB * 8192) + k * 256, lbo, sbo);
          mma_bf16_ss(tmem, da, db, idesc, (it | s | k) != 0);
        }
    mma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0);
    if (out_cycles) *out_cycles = clock64() - t0;
  }
  if (warp < 4) {
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) out[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}
}  // namespace

extern "C" int rsn_probe_umma_mnmajor_cm(const void* u_blocks, const void* v_blocks, int64_t n_blocks, int64_t lbo, int64_t sbo,
                                         float* out, int64_t iters, long long* out_cycles, cudaStream_t stream) {
  RSN_ARG(n_blocks >= 1 && n_blocks <= 4, "rsn_probe_umma_mnmajor_cm: n_blocks in [1,4]");
  size_t smem = (size_t)2 * (2 + n_blocks) * 8192 + 1024;
  RSN_CUDA(cudaFuncSetAttribute(probe_mnmajor_cm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_mnmajor_cm_kernel<<<1, 160, smem, stream>>>((const uint8_t*)u_blocks, (const uint8_t*)v_blocks, (int)n_blocks, (int)lbo,
                                                    (int)sbo, out, (int)iters, out_cycles);
  RSN_LAUNCH_CHECK("probe_mnmajor_cm_kernel");
  return 0;
}

// ---- A-from-TMEM probe: D[128, N] = X[128, K] * W[N, K]^T with X written into TMEM by the four epilogue-style warps
// (tcgen05.st, two bf16 per 32-bit column) and read by tcgen05.mma as its A operand; W from shared memory.
// `iters` > 0 additionally times `iters` back-to-back MMAs of the same form (cycles -> out_cycles).
namespace {
__global__ void __launch_bounds__(160, 1) probe_ts_kernel(const uint8_t* __restrict__ x_blocks,
                                                          const uint8_t* __restrict__ w_blocks, int N, int KB,
                                                          float* __restrict__ out, int iters, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_load, bar_mma, bar_a;
  __shared__ uint32_t tmem_base_slot;
  uint8_t* sx = smem;
  uint8_t* sw = smem + (size_t)KB * 16384;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w_block_bytes = (uint32_t)N * 128u;
  constexpr uint32_t A_COL = 256;

  if (warp == 4) {
    if (lane == 0) {
      mbar_init(&bar_load, 1);
      mbar_init(&bar_mma, 1);
      mbar_init(&bar_a, 128);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_base_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;

  if (warp == 4 && lane == 0) {
    mbar_expect_tx(&bar_load, (uint32_t)KB * (16384u + w_block_bytes));
    for (int kb = 0; kb < KB; ++kb) {
      bulk_g2s(sx + (size_t)kb * 16384, x_blocks + (size_t)kb * 16384, 16384, &bar_load);
      bulk_g2s(sw + (size_t)kb * w_block_bytes, w_blocks + (size_t)kb * w_block_bytes, w_block_bytes, &bar_load);
    }
  }
  if (warp < 4) {
    // X rows -> TMEM: column A_COL + kb*32 + j holds elements (64 kb + 2j, 64 kb + 2j + 1) of this thread's row
    mbar_wait(&bar_load, 0);
    const int row = warp * 32 + lane;
    for (int kb = 0; kb < KB; ++kb) {
      uint32_t v[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 t = *reinterpret_cast<const uint4*>(sx + (size_t)kb * 16384 + row * 128 + ((c ^ (row & 7)) << 4));
        v[c * 4 + 0] = t.x, v[c * 4 + 1] = t.y, v[c * 4 + 2] = t.z, v[c * 4 + 3] = t.w;
      }
      tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + A_COL + kb * 32, v);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(&bar_a);
  }
  if (warp == 4 && lane == 0) {
    mbar_wait(&bar_load, 0);
    mbar_wait(&bar_a, 0);
    tc_fence_after();
    const uint32_t idesc = instr_desc_bf16(128, N, 0, 0);
    constexpr uint32_t HI = desc_hi_sw128(1024);
    for (int kb = 0; kb < KB; ++kb)
      for (int k = 0; k < 4; ++k)
        mma_bf16_ts_lo(tmem, tmem + A_COL + kb * 32 + k * 8, desc_lo(smem_u32(sw + (size_t)kb * w_block_bytes), 16) + k * 2,
                       HI, idesc, (kb | k) != 0);
    mma_commit(&bar_mma);
  }
  if (warp < 4) {
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) out[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (iters > 0 && warp == 4 && lane == 0) {
    tc_fence_after();
    const uint32_t idesc = instr_desc_bf16(128, N, 0, 0);
    constexpr uint32_t HI = desc_hi_sw128(1024);
    const uint32_t b_lo = desc_lo(smem_u32(sw), 16);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_bf16_ts_lo(tmem, tmem + A_COL + k * 8, b_lo + k * 2, HI, idesc, 1u);
    }
    mma_commit(&bar_mma);
    mbar_wait(&bar_mma, 1);
    const long long t1 = clock64();
    out_cycles[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}
}  // namespace

extern "C" int rsn_probe_umma_ts(const void* x_blocks, const void* w_blocks, int64_t n_out, int64_t k_blocks, float* out,
                                 int64_t iters, int64_t* cycles_out, cudaStream_t stream) {
  RSN_ARG(n_out >= 16 && n_out <= 256 && n_out % 16 == 0, "rsn_probe_umma_ts: n_out in [16,256], multiple of 16");
  RSN_ARG(k_blocks >= 1 && k_blocks <= 4, "rsn_probe_umma_ts: k_blocks in [1,4]");
  RSN_ARG(iters >= 0 && iters % 4 == 0 && (iters == 0 || cycles_out), "rsn_probe_umma_ts: bad iters");
  size_t smem = (size_t)k_blocks * (16384 + n_out * 128) + 1024;
  RSN_CUDA(cudaFuncSetAttribute(probe_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_ts_kernel<<<1, 160, smem, stream>>>((const uint8_t*)x_blocks, (const uint8_t*)w_blocks, (int)n_out, (int)k_blocks,
                                            out, (int)iters, (long long*)cycles_out);
  RSN_LAUNCH_CHECK("probe_ts_kernel");
  return 0;
}

// ---- TMEM read / write throughput as the epilogue warps see it: `n_warps` warps (1..8; warp w touches lane quarter
// w % 4) each move 64 columns x 32 lanes per iteration.  mode 0: 2 x tcgen05.ld.x32 + wait; 1: 4 x ld.x16 + wait;
// 2: mode 0 followed by one tcgen05.st.x32 + wait::st; 3: mode 0 with the wait only every 4th iteration.
// mma_iters > 0: a ninth warp issues that many back-to-back M128 x N256 x K16 MMAs into the other 256 columns meanwhile
// (cycles_out[8] = its time): what the field kernels' epilogue sees while the next layer's MMAs run.
namespace {
__global__ void __launch_bounds__(288, 1) probe_tmem_kernel(int n_warps, int mode, int iters, int mma_iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_base_slot;
  __shared__ uint64_t bar_mma;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 6 * 16384 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (warp == 8) {
    if (lane == 0) {
      mbar_init(&bar_mma, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_base_slot, 512);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  if (warp < n_warps) {
    // the readers sweep columns 0..255; a concurrent MMA stream (mma_iters > 0) accumulates into columns 256..511
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t col = (uint32_t)((i + (warp >> 2) * 2) & 3) * 64u;
      if (mode == 1) {
        uint32_t v[4][16];
#pragma unroll
        for (int j = 0; j < 4; ++j) tmem_ld16(tl + col + j * 16, v[j]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += v[j][0] ^ v[j][15];
      } else {
        uint32_t v[2][32];
        tmem_ld32(tl + col, v[0]);
        tmem_ld32(tl + col + 32, v[1]);
        if (mode != 3 || (i & 3) == 3) tmem_ld_wait();
        acc += v[0][0] ^ v[1][31];
        if (mode == 2) {
          tmem_st32(tl + col, v[1]);
          tmem_st_wait();
        }
      }
    }
    tmem_ld_wait();
    const long long t1 = clock64();
    if (lane == 0) out[warp] = (t1 - t0) + (acc == 0x12345u ? 1 : 0);
  }
  if (warp == 8 && lane == 0 && mma_iters > 0) {
    const uint32_t idesc = instr_desc_bf16(128, 256, 0, 0);
    const uint32_t a_lo = desc_lo(smem_u32(smem), 16), b_lo = desc_lo(smem_u32(smem) + 2 * 16384, 16);
    const long long t0 = clock64();
    for (int i = 0; i < mma_iters; i += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_bf16_ss_lo(tmem + 256, a_lo + k * 2, b_lo + k * 2, desc_hi_sw128(1024), idesc, 1u);
    }
    mma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0);
    out[8] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}
}  // namespace

extern "C" int rsn_probe_tmem_rate(int64_t n_warps, int64_t mode, int64_t iters, int64_t mma_iters, int64_t* cycles_out,
                                   cudaStream_t stream) {
  RSN_ARG(n_warps >= 1 && n_warps <= 8 && mode >= 0 && mode <= 3 && iters > 0 && cycles_out, "rsn_probe_tmem_rate: bad arguments");
  RSN_ARG(mma_iters >= 0 && mma_iters % 4 == 0, "rsn_probe_tmem_rate: mma_iters must be a multiple of 4");
  const size_t smem = 6 * 16384 + 1024;
  RSN_CUDA(cudaFuncSetAttribute(probe_tmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_tmem_kernel<<<1, 288, smem, stream>>>((int)n_warps, (int)mode, (int)iters, (int)mma_iters, (long long*)cycles_out);
  RSN_LAUNCH_CHECK("probe_tmem_kernel");
  return 0;
}

// ---- cost of the field kernels' epilogue step by step: four warps, each iteration converts one 64-column group of
// its 32 rows.  `steps` bits: 1 = bias add + ReLU + bf16 pack, 2 = 8 x st.shared.v4 into a swizzled block,
// 4 = fence.proxy.async, 8 = tcgen05.fence::before_thread_sync + mbarrier.arrive by every thread, 16 = bias from
// constant memory (else a register constant), 32 = tcgen05.st of the packed row + wait::st, 64 = one arrive per warp
// instead of per thread.  mma_iters > 0: concurrent SS MMA stream as in probe_tmem_kernel.
namespace {
__constant__ float4 c_probe_bias[64];
__global__ void __launch_bounds__(288, 1) probe_epi_kernel(int steps, int iters, int mma_iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_base_slot;
  __shared__ uint64_t bar_mma, bar_grp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 10 * 16384 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (warp == 8) {
    if (lane == 0) {
      mbar_init(&bar_mma, 1);
      mbar_init(&bar_grp, (steps & 64) ? 4 : 128);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_base_slot, 512);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  if (warp < 4) {
    const int row = warp * 32 + lane;
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t s_act = smem_u32(smem) + 6 * 16384;
    uint32_t sink = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int g = it & 3;
      uint32_t v[2][32];
      tmem_ld32(tl + g * 64, v[0]);
      tmem_ld32(tl + g * 64 + 32, v[1]);
      float4 b[16];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        b[i] = (steps & 16) ? c_probe_bias[g * 16 + i] : make_float4(0.1f, 0.2f, 0.3f, 0.4f);
      tmem_ld_wait();
      uint32_t a[32];
      const uint32_t row_saddr = s_act + (uint32_t)g * 16384u + (uint32_t)row * 128u;
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[4];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float4 bb = b[h * 8 + c * 2 + q];
            float x0 = __uint_as_float(v[h][c * 8 + q * 4 + 0]), x1 = __uint_as_float(v[h][c * 8 + q * 4 + 1]);
            float x2 = __uint_as_float(v[h][c * 8 + q * 4 + 2]), x3 = __uint_as_float(v[h][c * 8 + q * 4 + 3]);
            if (steps & 1) {
              x0 += bb.x, x1 += bb.y, x2 += bb.z, x3 += bb.w;
              asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(pk[q * 2 + 0]) : "f"(x1), "f"(x0));
              asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(pk[q * 2 + 1]) : "f"(x3), "f"(x2));
            } else {
              pk[q * 2 + 0] = __float_as_uint(x0) ^ __float_as_uint(x1);
              pk[q * 2 + 1] = __float_as_uint(x2) ^ __float_as_uint(x3);
            }
          }
          const int chunk = h * 4 + c;
          if (steps & 2)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_saddr + (uint32_t)((chunk ^ (row & 7)) << 4)),
                         "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                         : "memory");
#pragma unroll
          for (int j = 0; j < 4; ++j) a[chunk * 4 + j] = pk[j];
        }
      if (steps & 32) {
        tmem_st32(tl + g * 32, a);
        tmem_st_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) sink ^= a[j];
      }
      if (steps & 4) fence_proxy_async();
      if (steps & 8) {
        tc_fence_before();
        if (steps & 64) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_grp);
        } else {
          mbar_arrive(&bar_grp);
        }
      }
    }
    const long long t1 = clock64();
    if (lane == 0) out[warp] = (t1 - t0) + (sink == 0x12345u ? 1 : 0);
  }
  if (warp == 8 && lane == 0 && mma_iters > 0) {
    const uint32_t idesc = instr_desc_bf16(128, 256, 0, 0);
    const uint32_t a_lo = desc_lo(smem_u32(smem), 16), b_lo = desc_lo(smem_u32(smem) + 2 * 16384, 16);
    const long long t0 = clock64();
    for (int i = 0; i < mma_iters; i += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_bf16_ss_lo(tmem + 256, a_lo + k * 2, b_lo + k * 2, desc_hi_sw128(1024), idesc, 1u);
    }
    mma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0);
    out[8] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}
}  // namespace

extern "C" int rsn_probe_epilogue(int64_t steps, int64_t iters, int64_t mma_iters, int64_t* cycles_out, cudaStream_t stream) {
  RSN_ARG(steps >= 0 && steps < 128 && iters > 0 && mma_iters >= 0 && mma_iters % 4 == 0 && cycles_out,
          "rsn_probe_epilogue: bad arguments");
  const size_t smem = 10 * 16384 + 1024;
  RSN_CUDA(cudaFuncSetAttribute(probe_epi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_epi_kernel<<<1, 288, smem, stream>>>((int)steps, (int)iters, (int)mma_iters, (long long*)cycles_out);
  RSN_LAUNCH_CHECK("probe_epi_kernel");
  return 0;
}

// ---- tcgen05.mma issue-rate probe: cycles per M128 x N x K16 bf16 MMA for the four operand major-ness
// combinations (operands: whatever is in shared memory; only the timing matters).  `cheap` selects the issue
// path: 0 = both 64-bit descriptors rebuilt per MMA, 1 = constant high word + 32-bit add (mma_bf16_ss_lo).
namespace {
__global__ void __launch_bounds__(160, 1) probe_rate_kernel(int a_major, int b_major, int cheap, int N, int iters,
                                                            long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_mma, bar_scratch;
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 6 * 16384 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  __syncthreads();
  if (warp == 4) {
    if (lane == 0) {
      mbar_init(&bar_mma, 1);
      mbar_init(&bar_scratch, 1);   // phases just keep completing; nobody waits on it
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_base_slot, 512);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = instr_desc_bf16(128, N, a_major, b_major);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 2 * 16384;
    const long long t0 = clock64();
    if (!cheap) {
      for (int i = 0; i < iters; ++i) {
        const int k = i & 3;
        const uint64_t da = a_major ? smem_desc_sw128(sa + k * 2048, 16384, 1024) : smem_desc_sw128(sa + k * 32, 16, 1024);
        const uint64_t db = b_major ? smem_desc_sw128(sb + k * 2048, 16384, 1024) : smem_desc_sw128(sb + k * 32, 16, 1024);
        mma_bf16_ss(tmem, da, db, idesc, i != 0);
      }
    } else {
      // cheap >> 1: bits 0-3 = commit every 4*c MMAs to a scratch barrier (0 = never), bits 4-7 = switch between
      // the two 256-column accumulators every 4*s MMAs (0 = never)
      const int commit_every = ((cheap >> 1) & 15) * 4, switch_every = ((cheap >> 5) & 15) * 4;
      const uint32_t a_lo = desc_lo(sa, a_major ? 16384 : 16), b_lo = desc_lo(sb, b_major ? 16384 : 16);
      const uint32_t as = a_major ? 128 : 2, bs = b_major ? 128 : 2;
      for (int i = 0; i < iters; i += 4) {
        const uint32_t tm = tmem + ((switch_every && ((i / switch_every) & 1)) ? 256u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_bf16_ss_lo(tm, a_lo + k * as, b_lo + k * bs, desc_hi_sw128(1024), idesc, i >= 2 * (switch_every ? switch_every : 4));
        if (commit_every && ((i + 4) % commit_every) == 0) mma_commit(&bar_scratch);
      }
    }
    mma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}
}  // namespace

extern "C" int rsn_probe_umma_rate(int a_major, int b_major, int64_t n, int64_t iters, int64_t* cycles_out,
                                   cudaStream_t stream) {
  RSN_ARG(n >= 16 && n <= 256 && n % 16 == 0 && iters > 0 && iters % 4 == 0, "rsn_probe_umma_rate: bad arguments");
  size_t smem = 6 * 16384 + 1024;
  RSN_CUDA(cudaFuncSetAttribute(probe_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // a_major bit 1: cheap issue path; bits 8-15: number of CTAs (CTA 0 reports); bits 16-19 / 20-23: commit / switch
  // the accumulator every 4*x MMAs (cheap path only)
  const int grid = ((a_major >> 8) & 255) ? ((a_major >> 8) & 255) : 1;
  const int cheap = ((a_major >> 1) & 1) | (((a_major >> 16) & 255) << 1);
  probe_rate_kernel<<<grid, 160, smem, stream>>>(a_major & 1, b_major & 1, cheap, (int)n, (int)iters,
                                                 (long long*)cycles_out);
  RSN_LAUNCH_CHECK("probe_rate_kernel");
  return 0;
}

// ---- CTA-pair probe: D[256, N] = X[256, K] * W[N, K]^T with tcgen05.mma.cta_group::2.  CTA r of the pair stages its
// own 128 rows of X and rows [r N/2, (r+1) N/2) of W; the leader issues; both read their 128 x N result from TMEM.
namespace {
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160, 1)
    probe_2cta_kernel(const uint8_t* __restrict__ x_blocks, const uint8_t* __restrict__ w_blocks, int N, int KB,
                      float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_load, bar_peer, bar_mma;
  __shared__ uint32_t tmem_base_slot;
  const uint32_t rank = cluster_ctarank();
  uint8_t* sx = smem;
  uint8_t* sw = smem + (size_t)KB * 16384;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t half_bytes = (uint32_t)(N / 2) * 128u;

  if (warp == 4) {
    if (lane == 0) {
      mbar_init(&bar_load, 1);
      mbar_init(&bar_peer, 1);
      mbar_init(&bar_mma, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2cta(&tmem_base_slot, 512);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;

  if (warp == 4 && lane == 0) {
    mbar_expect_tx(&bar_load, (uint32_t)KB * (16384u + half_bytes));
    for (int kb = 0; kb < KB; ++kb) {
      bulk_g2s(sx + (size_t)kb * 16384, x_blocks + ((size_t)rank * KB + kb) * 16384, 16384, &bar_load);
      bulk_g2s(sw + (size_t)kb * half_bytes, w_blocks + (size_t)kb * N * 128 + (size_t)rank * half_bytes, half_bytes, &bar_load);
    }
    mbar_wait(&bar_load, 0);
    if (rank == 1) {
      mbar_arrive_remote(&bar_peer, 0);      // my operands are in my shared memory
    } else {
      mbar_wait_cluster(&bar_peer, 0);
      tc_fence_after();
      const uint32_t idesc = instr_desc_bf16(256, N, 0, 0);
      for (int kb = 0; kb < KB; ++kb)
        for (int k = 0; k < 4; ++k)
          mma_bf16_ss_lo_2cta(tmem, desc_lo(smem_u32(sx + (size_t)kb * 16384), 16) + 2 * k,
                              desc_lo(smem_u32(sw + (size_t)kb * half_bytes), 16) + 2 * k, desc_hi_sw128(1024), idesc,
                              (kb | k) != 0);
      mma_commit_2cta(&bar_mma);
    }
  }
  if (warp < 4) {
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) out[((size_t)rank * 128 + row) * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 4) tmem_dealloc_2cta(tmem, 512);
}
}  // namespace

extern "C" int rsn_probe_umma_2cta(const void* x_blocks, const void* w_blocks, int64_t n_out, int64_t k_blocks, float* out,
                                   cudaStream_t stream) {
  RSN_ARG(n_out >= 16 && n_out <= 256 && n_out % 16 == 0, "rsn_probe_umma_2cta: n_out in [16,256], multiple of 16");
  RSN_ARG(k_blocks >= 1 && k_blocks <= 4, "rsn_probe_umma_2cta: k_blocks in [1,4]");
  size_t smem = (size_t)k_blocks * (16384 + n_out * 64) + 1024;
  RSN_CUDA(cudaFuncSetAttribute(probe_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_2cta_kernel<<<2, 160, smem, stream>>>((const uint8_t*)x_blocks, (const uint8_t*)w_blocks, (int)n_out,
                                              (int)k_blocks, out);
  RSN_LAUNCH_CHECK("probe_2cta_kernel");
  return 0;
}

// ---- issue-rate probe for the CTA pair: cycles per M256 x N x K16 cta_group::2 MMA (leader's clock)
namespace {
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160, 1) probe_rate2_kernel(int N, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 6 * 16384 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  __syncthreads();
  if (warp == 4) {
    if (lane == 0) {
      mbar_init(&bar_mma, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2cta(&tmem_base_slot, 512);
  }
  fence_proxy_async();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  if (warp == 4 && lane == 0 && rank == 0) {
    const uint32_t idesc = instr_desc_bf16(256, N, 0, 0);
    const uint32_t a_lo = desc_lo(smem_u32(smem), 16), b_lo = desc_lo(smem_u32(smem) + 2 * 16384, 16);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma_bf16_ss_lo_2cta(tmem, a_lo + 2 * k, b_lo + 2 * k, desc_hi_sw128(1024), idesc, (i | k) != 0);
    }
    mma_commit_2cta(&bar_mma);
    mbar_wait(&bar_mma, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  if (rank == 1 && warp == 4 && lane == 0) mbar_wait(&bar_mma, 0);
  tc_fence_before();
  cluster_sync_all();
  if (warp == 4) tmem_dealloc_2cta(tmem, 512);
}
}  // namespace

extern "C" int rsn_probe_umma_rate_2cta(int64_t n, int64_t iters, int64_t n_pairs, int64_t* cycles_out, cudaStream_t stream) {
  RSN_ARG(n >= 16 && n <= 256 && n % 16 == 0 && iters > 0 && iters % 4 == 0 && n_pairs >= 1, "rsn_probe_umma_rate_2cta: bad arguments");
  size_t smem = 6 * 16384 + 1024;
  RSN_CUDA(cudaFuncSetAttribute(probe_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_rate2_kernel<<<(int)(2 * n_pairs), 160, smem, stream>>>((int)n, (int)iters, (long long*)cycles_out);
  RSN_LAUNCH_CHECK("probe_rate2_kernel");
  return 0;
}
