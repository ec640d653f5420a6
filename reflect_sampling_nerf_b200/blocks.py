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
