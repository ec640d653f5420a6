// Packs the field's fp32 parameters into the operand images the fused kernels stream (csrc/field_layout.cuh):
// the forward weight blob (MMA consumption order), the transposed blob of the dgrad chains, the bias vector and
// the bf16 density-head row.  One launch, table driven; replaces ~200 small tensor ops per optimizer step.
//
// Parameter order of the `params` array (device pointers, fp32, row-major [out, in] as nn.Linear stores them;
// names of reflect_sampling_nerf_field.py:54-86):
//    0..7   mlp_base.layers.{0..7}.weight        8..15  mlp_base.layers.{0..7}.bias
//   16/17   field_output_bottleneck.net.{weight,bias}      18/19  mlp_mid.layers.0.{weight,bias}
//   20/21   field_output_mid.net.{weight,bias}
//   22/23   field_output_density.net   24/25 field_output_normals.net   26/27 field_output_roughness.net
//   28/29   field_output_diff.net      30/31 field_output_tint.net      ({weight,bias} each)
#include "rsn_common.cuh"
#include "umma.cuh"
#include "field_layout.cuh"

extern "C" int rsn_field_wgrad_layout(int64_t* host_offsets, int64_t* host_shapes, int64_t* total_floats);

namespace {

using namespace rsnf;

constexpr int N_PARAMS = 32;
constexpr int MAX_PIECES = 48;

// dst image [k_blocks][dst_rows][64] bf16 (swizzled) at byte offset dst_off of blob `which`;
// element (n, k) = M[n][k] (or M[k][n] when transposed) of the source matrix M = rows [0, ..) x columns
// [col0, col0 + ..) of parameter `param` (leading dimension ld), zero outside n < n_valid, k < k_valid.
// param < 0: M = the concatenated small heads [16][256] (density, normals 3, roughness, diff 3, tint 3, 5 zero rows).
struct Piece {
  int which;            // 0 forward blob, 1 transposed blob
  uint32_t dst_off;
  int dst_rows, k_blocks;
  int param, ld, col0, transposed;
  int n_valid, k_valid;
  int chunk0;           // first 16-byte chunk of this piece in the global chunk numbering
};
struct PackParams {
  const float* params[N_PARAMS];
  uint8_t* blob[2];
  float* bias;
  __nv_bfloat16* wd;
  int n_pieces, n_chunks;
  Piece pieces[MAX_PIECES];
};

// row h (0..15) of the concatenated heads matrix, or NULL for the zero rows
__device__ __forceinline__ const float* head_row(const PackParams& p, int h) {
  if (h == 0) return p.params[22];
  if (h < 4) return p.params[24] + (size_t)(h - 1) * 256;
  if (h == 4) return p.params[26];
  if (h < 8) return p.params[28] + (size_t)(h - 5) * 256;
  if (h < 11) return p.params[30] + (size_t)(h - 8) * 256;
  return nullptr;
}

__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ PackParams p) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  // ---- bias vector + density row (first N_BIAS + 256 threads)
  if (tid < N_BIAS) {
    float v = 0.f;
    if (tid < BIAS_BOTT) v = p.params[8 + tid / 256][tid % 256];
    else if (tid < BIAS_HEAD) v = p.params[17][tid - BIAS_BOTT];
    else if (tid < BIAS_MID) {
      const int h = tid - BIAS_HEAD;   // density 0 | normals 1-3 | roughness 4 | diff 5-7 | tint 8-10
      if (h == 0) v = p.params[23][0];
      else if (h < 4) v = p.params[25][h - 1];
      else if (h == 4) v = p.params[27][0];
      else if (h < 8) v = p.params[29][h - 5];
      else if (h < 11) v = p.params[31][h - 8];
    } else if (tid < BIAS_RGB) v = p.params[19][tid - BIAS_MID];
    else if (tid < BIAS_RGB + 3) v = p.params[21][tid - BIAS_RGB];
    p.bias[tid] = v;
  } else if (tid < N_BIAS + 256) {
    p.wd[tid - N_BIAS] = __float2bfloat16(p.params[22][tid - N_BIAS]);
  }
  // ---- operand images: one 16-byte chunk (8 consecutive k of one row n) per thread
  for (int c = tid; c < p.n_chunks; c += gridDim.x * blockDim.x) {
    int pi = 0;
    while (pi + 1 < p.n_pieces && p.pieces[pi + 1].chunk0 <= c) ++pi;
    const Piece& pc = p.pieces[pi];
    const int local = c - pc.chunk0;
    const int per_block = pc.dst_rows * 8;
    const int kb = local / per_block, n = (local % per_block) / 8, ch = local % 8;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = kb * 64 + ch * 8 + 2 * j + e;
        float x = 0.f;
        if (n < pc.n_valid && k < pc.k_valid) {
          const int r = pc.transposed ? k : n, cc = pc.transposed ? n : k;   // (row, column) of the source matrix
          if (pc.param >= 0) {
            x = p.params[pc.param][(size_t)r * pc.ld + pc.col0 + cc];
          } else {
            const float* hr = head_row(p, r);
            x = hr ? hr[cc] : 0.f;
          }
        }
        v[e] = x;
      }
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[j]) : "f"(v[1]), "f"(v[0]));
    }
    uint8_t* dst = p.blob[pc.which] + pc.dst_off + (size_t)kb * pc.dst_rows * 128 + (size_t)n * 128 + ((ch ^ (n & 7)) << 4);
    *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

struct Builder {
  PackParams p;
  int chunks = 0;
  void add(int which, uint32_t dst_off, int dst_rows, int k_blocks, int param, int ld, int col0, int transposed,
           int n_valid, int k_valid) {
    Piece& q = p.pieces[p.n_pieces++];
    q = Piece{which, dst_off, dst_rows, k_blocks, param, ld, col0, transposed, n_valid, k_valid, chunks};
    chunks += dst_rows * 8 * k_blocks;
  }
};

}  // namespace

extern "C" int rsn_pack_field(const float* const* params, void* wblob, void* wblob_t, float* bias, void* wd_bf16,
                              cudaStream_t stream) {
  RSN_ARG(params && wblob && wblob_t && bias && wd_bf16, "rsn_pack_field: null pointer");
  for (int i = 0; i < N_PARAMS; ++i) RSN_ARG(params[i] != nullptr, "rsn_pack_field: params[%d] is null", i);
  RSN_ARG(((uintptr_t)wblob & 15) == 0 && ((uintptr_t)wblob_t & 15) == 0, "rsn_pack_field: blobs must be 16-byte aligned");
  Builder b;
  b.p = PackParams{};
  for (int i = 0; i < N_PARAMS; ++i) b.p.params[i] = params[i];
  b.p.blob[0] = (uint8_t*)wblob;
  b.p.blob[1] = (uint8_t*)wblob_t;
  b.p.bias = bias;
  b.p.wd = (__nv_bfloat16*)wd_bf16;
  // ---------------- forward blob (chunk order of field_fwd.cu): B[n = output feature][k = input feature]
  uint32_t off = 0;
  auto fwd = [&](int rows, int kb, int param, int ld, int col0, int n_valid, int k_valid) {
    b.add(0, off, rows, kb, param, ld, col0, 0, n_valid, k_valid);
    off += (uint32_t)rows * 128u * kb;
  };
  for (int l = 0; l < 8; ++l) {
    if (l == 0) {
      fwd(256, 2, 0, 99, 0, 256, 99);
    } else if (l == 4) {
      fwd(256, 2, 4, 355, 0, 256, 99);      // encoding part
      fwd(256, 4, 4, 355, 99, 256, 256);    // hidden part
    } else {
      fwd(256, 4, l, 256, 0, 256, 256);
    }
  }
  fwd(256, 4, 16, 256, 0, 256, 256);        // bottleneck
  fwd(16, 4, -1, 256, 0, 16, 256);          // heads
  fwd(128, 4, 18, 290, 34, 128, 256);       // mid: bottleneck part (two chunks of two K-blocks)
  fwd(128, 1, 18, 290, 0, 128, 34);         // mid: IDE part
  fwd(16, 2, 20, 128, 0, 3, 128);           // rgb
  RSN_ARG(off == FWD_BLOB_BYTES, "rsn_pack_field: forward layout mismatch (%u)", off);
  // ---------------- transposed blob (csrc/field_layout.cuh BT_*): B[n = input feature][k = output feature]
  b.add(1, BT_RGB, 128, 1, 20, 128, 0, 1, 128, 3);
  b.add(1, BT_MID, 256, 2, 18, 290, 34, 1, 256, 128);
  b.add(1, BT_BOTT, 256, 4, 16, 256, 0, 1, 256, 256);
  b.add(1, BT_HEADS, 256, 1, -1, 256, 0, 1, 256, 16);
  for (int l = 1; l < 8; ++l) b.add(1, BT_L(l), 256, 4, l, l == 4 ? 355 : 256, l == 4 ? 99 : 0, 1, 256, 256);
  b.add(1, BT_L4E, 128, 4, 4, 355, 0, 1, 99, 256);
  b.add(1, BT_L0, 128, 4, 0, 99, 0, 1, 99, 256);
  RSN_ARG(b.p.n_pieces <= MAX_PIECES, "rsn_pack_field: piece table overflow");
  b.p.n_chunks = b.chunks;
  const int threads = 256, blocks = (b.chunks + threads - 1) / threads;
  pack_kernel<<<blocks, threads, 0, stream>>>(b.p);
  RSN_LAUNCH_CHECK("pack_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// Gradient blob of rsn_field_wgrad -> flat fp32 gradient vector in the parameter order of rsn_pack_field
// (weights row-major [out, in]); the Python side hands out views of the flat vector as the parameters' .grad.
// One launch instead of ~70 slicing / concatenation / clone kernels per step.
namespace {

struct UnpackPiece {
  int dst0;          // first element of this piece in the flat vector
  int rows, cols;    // destination sub-matrix rows x cols (cols = 1 for bias pieces)
  int dst_ld, dst_col0;
  int src0, src_ld;  // blob float offset of element (0, 0) and the row stride in the blob
  int elem0;         // running element count (for the thread -> piece lookup)
};
constexpr int MAX_UNPACK = 48;
struct UnpackParams {
  const float* blob;
  float* flat;
  int n_pieces, n_elems;
  UnpackPiece pieces[MAX_UNPACK];
};

__global__ void __launch_bounds__(256) unpack_kernel(const __grid_constant__ UnpackParams p) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < p.n_elems; e += gridDim.x * blockDim.x) {
    int pi = 0;
    while (pi + 1 < p.n_pieces && p.pieces[pi + 1].elem0 <= e) ++pi;
    const UnpackPiece& q = p.pieces[pi];
    const int local = e - q.elem0;
    const int r = local / q.cols, c = local % q.cols;
    p.flat[q.dst0 + (size_t)r * q.dst_ld + q.dst_col0 + c] = p.blob[q.src0 + (size_t)r * q.src_ld + c];
  }
}

}  // namespace

// Flat-vector offsets (in floats) of the 32 parameters in rsn_pack_field order; returns the total length.
extern "C" int64_t rsn_field_flat_layout(int64_t* host_offsets32) {
  static const int rows[32] = {256, 256, 256, 256, 256, 256, 256, 256, 256, 256, 256, 256, 256, 256, 256, 256,
                               256, 256, 128, 128, 3, 3, 1, 1, 3, 3, 1, 1, 3, 3, 3, 3};
  static const int cols[32] = {99, 256, 256, 256, 355, 256, 256, 256, 1, 1, 1, 1, 1, 1, 1, 1,
                               256, 1, 290, 1, 128, 1, 256, 1, 256, 1, 256, 1, 256, 1, 256, 1};
  int64_t off = 0;
  for (int i = 0; i < 32; ++i) {
    if (host_offsets32) host_offsets32[i] = off;
    off += (int64_t)rows[i] * cols[i];
  }
  return off;
}

extern "C" int rsn_unpack_grads(const float* grad_blob, float* flat_grads, cudaStream_t stream) {
  RSN_ARG(grad_blob && flat_grads, "rsn_unpack_grads: null pointer");
  int64_t offs[64], shapes[64], total = 0;
  const int n_jobs = rsn_field_wgrad_layout(offs, shapes, &total);
  RSN_ARG(n_jobs == 14, "rsn_unpack_grads: unexpected wgrad job table");
  int64_t po[32];
  rsn_field_flat_layout(po);
  UnpackParams u = {};
  u.blob = grad_blob;
  u.flat = flat_grads;
  int elems = 0;
  auto add = [&](int param, int rows, int cols, int dst_ld, int dst_col0, int64_t src0, int src_ld) {
    UnpackPiece& q = u.pieces[u.n_pieces++];
    q = UnpackPiece{(int)po[param], rows, cols, dst_ld, dst_col0, (int)src0, src_ld, elems};
    elems += rows * cols;
  };
  auto dw = [&](int j) { return offs[2 * j]; };
  auto db = [&](int j) { return offs[2 * j + 1]; };
  auto ld = [&](int j) { return (int)shapes[2 * j + 1]; };
  // base layers: job index of layer l (layer 4 = jobs 4 (enc part) + 5 (hidden part))
  const int base_job[8] = {0, 1, 2, 3, 5, 6, 7, 8};
  for (int l = 0; l < 8; ++l) {
    if (l == 0) add(0, 256, 99, 99, 0, dw(0), ld(0));
    else if (l == 4) {
      add(4, 256, 99, 355, 0, dw(4), ld(4));
      add(4, 256, 256, 355, 99, dw(5), ld(5));
    } else add(l, 256, 256, 256, 0, dw(base_job[l]), ld(base_job[l]));
    add(8 + l, 256, 1, 1, 0, db(base_job[l]), 1);
  }
  add(16, 256, 256, 256, 0, dw(9), ld(9));
  add(17, 256, 1, 1, 0, db(9), 1);                                     // bottleneck
  add(18, 128, 34, 290, 0, dw(13), ld(13));
  add(18, 128, 256, 290, 34, dw(12), ld(12));
  add(19, 128, 1, 1, 0, db(12), 1);                                    // mid
  add(20, 3, 128, 128, 0, dw(11), ld(11));
  add(21, 3, 1, 1, 0, db(10), 1);                                      // rgb (seed rows 0-2)
  const int head_param[5] = {22, 24, 26, 28, 30}, head_r0[5] = {0, 1, 4, 5, 8}, head_n[5] = {1, 3, 1, 3, 3};
  for (int h = 0; h < 5; ++h) {                                        // small heads: seed rows 16 + r0 ..
    add(head_param[h], head_n[h], 256, 256, 0, dw(10) + (int64_t)(16 + head_r0[h]) * ld(10), ld(10));
    add(head_param[h] + 1, head_n[h], 1, 1, 0, db(10) + 16 + head_r0[h], 1);
  }
  RSN_ARG(u.n_pieces <= MAX_UNPACK, "rsn_unpack_grads: piece table overflow");
  u.n_elems = elems;
  unpack_kernel<<<(elems + 255) / 256, 256, 0, stream>>>(u);
  RSN_LAUNCH_CHECK("unpack_kernel");
  return 0;
}
