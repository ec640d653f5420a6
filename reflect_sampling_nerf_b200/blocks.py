"""Host-side view of the operand "block image" used by the tcgen05 kernels (csrc/umma.cuh, block_off):
a block is [rows][64] bf16, 128 B per row, 128-byte swizzled -- the 16-byte chunk c of row r is stored
at chunk position (c ^ (r & 7)).  Blocks of one matrix are laid out K-block after K-block."""
from __future__ import annotations

import torch
from torch import Tensor


def _chunk_perm(rows: int, device) -> Tensor:
    r = torch.arange(rows, device=device)[:, None]
    c = torch.arange(8, device=device)[None, :]
    return c ^ (r & 7)                     # logical chunk c of row r -> physical chunk


def pack_blocks(mat: Tensor) -> Tensor:
    """[rows, K] (K multiple of 64, rows multiple of 8) bf16 -> uint8 image [K/64, rows, 128]."""
    rows, k = mat.shape
    assert k % 64 == 0 and rows % 8 == 0
    m = mat.to(torch.bfloat16).contiguous().view(rows, k // 64, 8, 8)        # [r, kb, chunk, 8 elems]
    phys = _chunk_perm(rows, mat.device)                                       # [r, 8]
    out = torch.empty_like(m)
    out.scatter_(2, phys[:, None, :, None].expand(rows, k // 64, 8, 8), m)
    return out.permute(1, 0, 2, 3).contiguous().view(torch.uint8).view(k // 64, rows, 128)


def unpack_blocks(img: Tensor, rows: int) -> Tensor:
    """Inverse of pack_blocks: uint8 image [KB, rows, 128] -> bf16 [rows, KB*64]."""
    kb = img.shape[0]
    m = img.contiguous().view(torch.bfloat16).view(kb, rows, 8, 8).permute(1, 0, 2, 3)
    phys = _chunk_perm(rows, img.device)
    out = torch.gather(m, 2, phys[:, None, :, None].expand(rows, kb, 8, 8))
    return out.reshape(rows, kb * 64)


# ---- the chunk-major image of the training stashes (csrc/field_layout.cuh, stash_chunk_off) --------------------------
# [128 points][64 features] bf16 block: the 16-byte chunk c (features 8c..8c+7) of point r at
# (r // 64) * 8192 + c * 1024 + (r % 64) * 16 -- two 64-point slabs, inside a slab one 1 KB run per chunk.
def pack_blocks_cm(mat: Tensor) -> Tensor:
    """[128 t, K] (K multiple of 64) bf16 -> uint8 image [t, K/64, 16384] of chunk-major blocks (one per tile and K block)."""
    rows, k = mat.shape
    assert k % 64 == 0 and rows % 128 == 0
    m = mat.to(torch.bfloat16).contiguous().view(rows // 128, 2, 64, k // 64, 8, 8)    # [tile, slab, r, kb, chunk, 8 elems]
    out = m.permute(0, 3, 1, 4, 2, 5).contiguous()                                     # [tile, kb, slab, chunk, r, 8]
    return out.view(torch.uint8).view(rows // 128, k // 64, 16384)


def unpack_blocks_cm(img: Tensor) -> Tensor:
    """Inverse of pack_blocks_cm: uint8 image [t, KB, 16384] -> bf16 [128 t, KB * 64]."""
    t, kb = img.shape[0], img.shape[1]
    m = img.contiguous().view(torch.bfloat16).view(t, kb, 2, 8, 64, 8).permute(0, 2, 4, 1, 3, 5)   # [tile, slab, r, kb, chunk, 8]
    return m.reshape(t * 128, kb * 64)
