// Layout constants shared by the fused field kernels (forward, dgrad chain, wgrad) and by the host-side
// packer (reflect_sampling_nerf_b200/packing.py mirrors these numbers; tests/test_packing.py checks them).
//
// The field is reflect_sampling_nerf_field.py:54-86:
//   mlp_base   8 x 256, input = 99-dim IPE, skip-concat [enc, h] at layer 4, ReLU after every layer
//   heads      density 1 | pred-normals 3 | roughness 1 | diff 3 | tint 3   (one N=16 GEMM, 11 used)
//   bottleneck 256 -> 256 (no activation)
//   mlp_mid    [IDE 34, bottleneck 256] -> 128, ReLU
//   mid rgb    128 -> 3, sigmoid
//
// GEMM "layers" of one 128-point tile, in issue order (li = 0..10):
//   0..7 base layers | 8 bottleneck (+ the heads GEMM riding on the same A operand) | 9 mid | 10 rgb
//
// Operands are "block images" (umma.cuh: block_off): [rows][64] bf16, 128-byte swizzled.
// Activations: rows = the 128 points of the tile.  Weights: rows = output features, one image per
// 64-wide K block, concatenated in the order the MMA issuer consumes them ("chunks").
#pragma once
#include <stdint.h>

namespace rsnf {

constexpr int TILE = 128;            // points per tile (= UMMA M)
constexpr int BLOCK_BYTES = 16384;   // one [128][64] bf16 activation block
constexpr int ENC_DIM = 99;          // IPE output (48 sin + 48 cos-like + xyz)
constexpr int ENC_KSTEPS_B1 = 3;     // second enc block: columns 64..111 (99 real + zero pad) = 3 K-steps
constexpr int IDE_DIM = 34;
constexpr int IDE_KSTEPS = 3;        // 34 -> 48 columns
constexpr int N_HEAD = 16;           // density, normals(3), roughness, diff(3), tint(3), 5 x pad
constexpr int HEAD_TMEM_COL = 240;   // heads accumulate in the OTHER accumulator buffer, columns 240..255

// ---- forward weight blob: chunk sizes in consumption order -----------------------------------------
constexpr int W_STAGE_BYTES = 32768;
constexpr int N_FWD_CHUNKS = 41;
constexpr uint32_t FWD_BLOB_BYTES = 36u * 32768u + 8192u + 32768u + 32768u + 16384u + 4096u;  // 1,273,856
__host__ __device__ constexpr uint32_t fwd_chunk_bytes(int c) {
  return c < 36 ? 32768u : c == 36 ? 8192u : c < 39 ? 32768u : c == 39 ? 16384u : 4096u;
}

// ---- transposed weight blob of the dgrad chains: B[n = input feature][k = output feature] K-block images ----
constexpr uint32_t BT_RGB = 0;                         // [128 mid-hidden][64 (3 used)]            16 KB
constexpr uint32_t BT_MID = 16384;                     // [256 bottleneck][128 mid-hidden]         2 x 32 KB
constexpr uint32_t BT_BOTT = BT_MID + 2 * 32768;       // [256 emb][256 bottleneck]                4 x 32 KB
constexpr uint32_t BT_HEADS = BT_BOTT + 4 * 32768;     // [256 emb][64 (11 used)]                  32 KB
constexpr uint32_t BT_LBASE = BT_HEADS + 32768;        // base layers 1..7 (layer 4: hidden part)  7 x 4 x 32 KB
__host__ __device__ constexpr uint32_t BT_L(int l) { return BT_LBASE + (uint32_t)(l - 1) * 131072u; }
constexpr uint32_t BT_L4E = BT_LBASE + 7 * 131072;     // [128 enc (99 used)][256]                 4 x 16 KB
constexpr uint32_t BT_L0 = BT_L4E + 65536;             // [128 enc (99 used)][256]                 4 x 16 KB
constexpr uint32_t BWD_BLOB_BYTES = BT_L0 + 65536;     // 1,294,336

// rows (output features) and K-blocks of forward chunk c: a CTA pair (cta_group::2) stages rows [r N/2, (r+1) N/2)
// of every K-block in CTA r
__host__ __device__ constexpr int fwd_chunk_rows(int c) { return c < 36 ? 256 : c == 36 ? 16 : c < 40 ? 128 : 16; }
__host__ __device__ constexpr int fwd_chunk_nkb(int c) { return c < 36 ? 1 : c == 36 ? 4 : c < 39 ? 2 : c == 39 ? 1 : 2; }

// ---- bias vector (fp32) -----------------------------------------------------------------------------
constexpr int BIAS_BASE = 0;          // 8 x 256
constexpr int BIAS_BOTT = 2048;       // 256
constexpr int BIAS_HEAD = 2304;       // 16
constexpr int BIAS_MID = 2320;        // 128
constexpr int BIAS_RGB = 2448;        // 16
constexpr int N_BIAS = 2464;

// ---- the two block images of the training stashes ----------------------------------------------------
// Every stash block holds [128 points][64 features] bf16 = 16 KB, in one of two images:
//  * "swizzled" (umma.cuh block_off): the K-major 128-byte-swizzled shared-memory operand image, bulk-stored as it sits
//    in shared memory.  Used for the blocks that exist in shared memory anyway: the IPE / IDE encodings (forward A
//    operands) -- STASH_ENC, STASH_IDE.
//  * "chunk-major": the 16-byte chunk c (features 8c..8c+7) of point r at
//        (r / 64) * 8192 + c * 1024 + (r % 64) * 16
//    i.e. two 64-point slabs (the wgrad's pipeline unit), inside a slab one 1 KB run per chunk with the points
//    contiguous.  Written straight from registers: the 32 lanes of a warp (32 consecutive points) store 512 contiguous
//    bytes per st.global.v4.  Read by the wgrad as a NO-swizzle MN-major UMMA operand (core matrix = 8 points x 16 B =
//    128 contiguous bytes; 1024 B between core matrices along the features, 128 B along the points).
//    Used for everything the epilogues produce: the hidden activations (STASH_H, STASH_MIDH) and every dY block.
constexpr int STASH_SLAB_ROWS = 64;
constexpr int STASH_CHUNK_STRIDE = STASH_SLAB_ROWS * 16;          // 1024
constexpr int STASH_SLAB_BYTES = 8 * STASH_CHUNK_STRIDE;          // 8192
__host__ __device__ constexpr uint32_t stash_chunk_off(int r, int c) {
  return (uint32_t)(r / STASH_SLAB_ROWS) * STASH_SLAB_BYTES + (uint32_t)c * STASH_CHUNK_STRIDE + (uint32_t)(r % STASH_SLAB_ROWS) * 16u;
}

// ---- training stash: per tile, STASH_BLOCKS activation block images of 16 KB (written by the forward) ----
constexpr int STASH_ENC = 0;     // 2 blocks: IPE (columns 99..127 zero)                         [swizzled image]
constexpr int STASH_H = 2;       // + 4*l + g : post-ReLU output of base layer l, 64-column group g   [chunk-major]
constexpr int STASH_BOTT = 34;   // 4 blocks: bottleneck (no activation)
constexpr int STASH_IDE = 38;    // 1 block: IDE (columns 34..63 zero)                            [swizzled image]
constexpr int STASH_MIDH = 39;   // 2 blocks: mid hidden (post-ReLU)                              [chunk-major]
constexpr int STASH_BLOCKS = 41;
// ... followed by the ReLU bit masks of the 8 hidden layers and the mid hidden layer: [9 layers][4 groups][128 rows]
// x 8 bytes; bit i (i < 16) of word w (w = 0, 1) of a row's entry = column 32 w + 2 i of the 64-column group is > 0,
// bit 16 + i = column 32 w + 2 i + 1.  The dgrad chains read these 288 B per point instead of the 4.4 KB of
// activations they mask with.
constexpr int MASK_LAYERS = 9;     // 0..7 = h_l, 8 = mid hidden (groups 0,1)
constexpr int STASH_MASK_OFF = STASH_BLOCKS * BLOCK_BYTES;
constexpr int STASH_MASK_BYTES = MASK_LAYERS * 4 * TILE * 8;               // 36,864
constexpr int STASH_TILE_BYTES = STASH_MASK_OFF + STASH_MASK_BYTES;        // 708,608 per 128-point tile
__host__ __device__ constexpr int mask_entry(int layer, int group, int row) { return ((layer * 4 + group) * TILE + row); }

// ---- dgrad stash: per tile, DY_BLOCKS chunk-major block images of the pre-activation gradients (written by the dgrad chain)
constexpr int DY_SEED = 0;       // 1 block: columns 0-15 d(rgb head pre-activation), 16-31 d(heads pre-activation)
constexpr int DY_MID = 1;        // 2 blocks: d(mid hidden pre-activation), 128 columns
constexpr int DY_BOTT = 3;       // 4 blocks: d(bottleneck)
constexpr int DY_H = 7;          // + 4*l + g : d(pre-activation of base layer l)
constexpr int DY_BLOCKS = 39;
constexpr bool DY_CHUNK_MAJOR = true;    // every dY block is a chunk-major image

// ---- per-point feature row written by the forward kernel ([P][16] fp32) ------------------------------
// 0-2 rgb = diff + tint*mid | 3-5 diff | 6-8 tint | 9-11 pred_normal | 12 sigmoid(rough) | 13 n.d
// 14 raw density (before softplus, without the 0.5 bias) | 15 softplus(rough)
constexpr int N_FEAT = 16;

}  // namespace rsnf
