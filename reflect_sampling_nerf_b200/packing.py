"""Packs the field's fp32 parameters (state-dict names of reflect_sampling_nerf_field.py:54-86) into the
operand images the fused kernels stream: a bf16 weight blob in MMA consumption order (csrc/field_layout.cuh)
and one fp32 bias vector.  The blob is derived state -- re-packed after every optimizer step, never
checkpointed.  Pure tensor ops: runs on whatever device the parameters live on.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
from torch import Tensor

from .blocks import pack_blocks

FWD_BLOB_BYTES = 36 * 32768 + 8192 + 32768 + 32768 + 16384 + 4096
N_BIAS = 2464
BIAS_BASE, BIAS_BOTT, BIAS_HEAD, BIAS_MID, BIAS_RGB = 0, 2048, 2304, 2320, 2448
HEAD_ROWS = {"density": (0, 1), "normals": (1, 4), "roughness": (4, 5), "diff": (5, 8), "tint": (8, 11)}
ENC_DIM, IDE_DIM = 99, 34


def _pad_cols(w: Tensor, k: int) -> Tensor:
    out = w.new_zeros(w.shape[0], k)
    out[:, : w.shape[1]] = w
    return out


def _pad_rows(w: Tensor, n: int) -> Tensor:
    out = w.new_zeros(n, w.shape[1])
    out[: w.shape[0]] = w
    return out


def head_matrix(sd: Dict[str, Tensor], prefix: str = "") -> Tuple[Tensor, Tensor]:
    """[16,256] weight / [16] bias of the concatenated small heads (rows: HEAD_ROWS, 5 zero rows)."""
    w = sd[prefix + "field_output_density.net.weight"]
    hw = w.new_zeros(16, 256)
    hb = w.new_zeros(16)
    for name, (lo, hi) in HEAD_ROWS.items():
        hw[lo:hi] = sd[f"{prefix}field_output_{name}.net.weight"]
        hb[lo:hi] = sd[f"{prefix}field_output_{name}.net.bias"]
    return hw, hb


def pack_field(sd: Dict[str, Tensor], prefix: str = "") -> Tuple[Tensor, Tensor]:
    """state dict of the field -> (uint8 blob [FWD_BLOB_BYTES], fp32 bias [N_BIAS])."""
    g = lambda k: sd[prefix + k].detach().float()  # noqa: E731
    parts = []
    bias = g("mlp_base.layers.0.bias").new_zeros(N_BIAS)
    for l in range(8):
        w = g(f"mlp_base.layers.{l}.weight")
        bias[BIAS_BASE + 256 * l: BIAS_BASE + 256 * (l + 1)] = g(f"mlp_base.layers.{l}.bias")
        if l == 0:
            parts.append(pack_blocks(_pad_cols(w, 128)))
        elif l == 4:  # [enc 99 | hidden 256] -> enc K-blocks first, then the hidden ones
            parts.append(pack_blocks(_pad_cols(w[:, :ENC_DIM], 128)))
            parts.append(pack_blocks(w[:, ENC_DIM:].contiguous()))
        else:
            parts.append(pack_blocks(w))
    parts.append(pack_blocks(g("field_output_bottleneck.net.weight")))
    bias[BIAS_BOTT: BIAS_BOTT + 256] = g("field_output_bottleneck.net.bias")
    hw, hb = head_matrix({k: v.detach().float() for k, v in sd.items()}, prefix)
    parts.append(pack_blocks(hw))
    bias[BIAS_HEAD: BIAS_HEAD + 16] = hb
    wm = g("mlp_mid.layers.0.weight")          # [128, 290] = [IDE 34 | bottleneck 256]
    parts.append(pack_blocks(wm[:, IDE_DIM:].contiguous()))
    parts.append(pack_blocks(_pad_cols(wm[:, :IDE_DIM], 64)))
    bias[BIAS_MID: BIAS_MID + 128] = g("mlp_mid.layers.0.bias")
    parts.append(pack_blocks(_pad_rows(g("field_output_mid.net.weight"), 16)))
    bias[BIAS_RGB: BIAS_RGB + 3] = g("field_output_mid.net.bias")
    blob = torch.cat([p.reshape(-1) for p in parts])
    assert blob.numel() == FWD_BLOB_BYTES, blob.numel()
    return blob.contiguous(), bias.contiguous()
