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


# ------------------------------------------------------------------------------------------ backward
BWD_BLOB_BYTES = 1294336
PARAM_NAMES = (
    [f"mlp_base.layers.{l}" for l in range(8)]
    + ["field_output_density.net", "field_output_bottleneck.net", "mlp_mid.layers.0", "field_output_mid.net",
       "field_output_normals.net", "field_output_roughness.net", "field_output_diff.net", "field_output_tint.net"])


def pack_field_t(sd: Dict[str, Tensor], prefix: str = "") -> Tuple[Tensor, Tensor]:
    """Transposed operand images of the dgrad chains (csrc/field_layout.cuh BT_*): for every Linear
    B[n = input feature][k = output feature].  Returns (uint8 blob [BWD_BLOB_BYTES], bf16 density row [256])."""
    g = lambda k: sd[prefix + k].detach().float()  # noqa: E731
    w4 = g("mlp_base.layers.4.weight")
    wm = g("mlp_mid.layers.0.weight")
    hw, _ = head_matrix({k: v.detach().float() for k, v in sd.items()}, prefix)
    parts = [
        pack_blocks(_pad_cols(g("field_output_mid.net.weight").T.contiguous(), 64)),            # BT_RGB   [128][64]
        pack_blocks(wm[:, IDE_DIM:].T.contiguous()),                                             # BT_MID   [256][128]
        pack_blocks(g("field_output_bottleneck.net.weight").T.contiguous()),                     # BT_BOTT  [256][256]
        pack_blocks(_pad_cols(hw.T.contiguous(), 64)),                                           # BT_HEADS [256][64]
    ]
    for l in range(1, 8):                                                                        # BT_L(1..7)
        w = w4[:, ENC_DIM:] if l == 4 else g(f"mlp_base.layers.{l}.weight")
        parts.append(pack_blocks(w.T.contiguous()))
    parts.append(pack_blocks(_pad_rows(w4[:, :ENC_DIM].T.contiguous(), 128)))                    # BT_L4E   [128][256]
    parts.append(pack_blocks(_pad_rows(g("mlp_base.layers.0.weight").T.contiguous(), 128)))      # BT_L0    [128][256]
    blob = torch.cat([p.reshape(-1) for p in parts])
    assert blob.numel() == BWD_BLOB_BYTES, blob.numel()
    wd = g("field_output_density.net.weight").reshape(256).to(torch.bfloat16).contiguous()
    return blob.contiguous(), wd


def unpack_grads(blob: Tensor, offsets, shapes) -> Dict[str, Tensor]:
    """Gradient blob of rsn_field_wgrad (regions: rsn_field_wgrad_layout) -> {parameter name: gradient}.
    Job order: csrc/field_wgrad.cu kJobs."""
    def region(j):
        o, (m, n) = offsets[2 * j], shapes[j]
        dw = blob[o: o + m * n].view(m, n)
        db = None if offsets[2 * j + 1] < 0 else blob[offsets[2 * j + 1]: offsets[2 * j + 1] + m]
        return dw, db

    out: Dict[str, Tensor] = {}
    base_jobs = {0: 0, 1: 1, 2: 2, 3: 3, 5: 6, 6: 7, 7: 8}
    for l, j in base_jobs.items():
        dw, db = region(j)
        out[f"mlp_base.layers.{l}.weight"] = dw[:, :ENC_DIM] if l == 0 else dw
        out[f"mlp_base.layers.{l}.bias"] = db
    dw_e, _ = region(4)
    dw_h, db4 = region(5)
    out["mlp_base.layers.4.weight"] = torch.cat([dw_e[:, :ENC_DIM], dw_h], dim=1)
    out["mlp_base.layers.4.bias"] = db4
    dw, db = region(9)
    out["field_output_bottleneck.net.weight"], out["field_output_bottleneck.net.bias"] = dw, db
    dw_hd, db_seed = region(10)       # seed block: rows 0-15 rgb head, rows 16-31 the small heads
    for name, (lo, hi) in HEAD_ROWS.items():
        out[f"field_output_{name}.net.weight"] = dw_hd[16 + lo: 16 + hi]
        out[f"field_output_{name}.net.bias"] = db_seed[16 + lo: 16 + hi]
    dw_rgb, _ = region(11)
    out["field_output_mid.net.weight"] = dw_rgb[:3]
    out["field_output_mid.net.bias"] = db_seed[:3]
    dw_mb, db_m = region(12)
    dw_mi, _ = region(13)
    out["mlp_mid.layers.0.weight"] = torch.cat([dw_mi[:, :IDE_DIM], dw_mb], dim=1)
    out["mlp_mid.layers.0.bias"] = db_m
    return out
