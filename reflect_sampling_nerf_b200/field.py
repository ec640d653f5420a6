"""Drop-in field: same parameters, module paths and state-dict keys as ReflectSamplingNeRFNerfField
(reflect_sampling_nerf_field.py:28-86), evaluated by the fused sm_100a kernels instead of eager PyTorch.

The reference evaluates the field through a sequence of small methods (get_blob -> contract -> get_density ->
get_pred_normals / get_roughness / get_diff / get_tint / get_mid, field.py:90-186) that all consume the same
samples; here the whole sequence is ONE kernel launch per pass (`evaluate_samples`), and get_inf_color
(field.py:190-201) is a second mode of the same kernel.  Parameters stay fp32 nn.Parameters owned by PyTorch
(optimizers, GradScaler, DDP and checkpoints keep working); the bf16 operand blob the kernels
stream is derived state, re-packed whenever a parameter changes.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn

from . import ops, packing


class _Layers(nn.Module):
    """nerfstudio MLP parameter container: `layers` = ModuleList of nn.Linear (keys `*.layers.{i}.weight`)."""

    def __init__(self, dims) -> None:
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(i, o) for i, o in dims])


class _Head(nn.Module):
    """nerfstudio FieldHead parameter container: `net` = nn.Linear (keys `*.net.weight`)."""

    def __init__(self, in_dim: int, out_dim: int) -> None:
        super().__init__()
        self.net = nn.Linear(in_dim, out_dim)


class ReflectSamplingNeRFNerfField(nn.Module):
    """B200-native field.  Constructor arguments mirror field.py:36-48; the architecture the kernels implement
    is the one the reference model instantiates (model.py:98-106): 99-dim IPE, 8x256 base MLP with a skip at
    layer 4, 34-dim IDE, one 128-wide mid layer."""

    ENC_DIM, IDE_DIM, WIDTH, MID_WIDTH = 99, 34, 256, 128

    def __init__(self, position_encoding=None, direction_encoding=None, base_mlp_num_layers: int = 8,
                 base_mlp_layer_width: int = 256, skip_connections: Tuple[int, ...] = (4,),
                 head_mlp_num_layers: int = 1, head_mlp_layer_width: int = 128, spatial_distortion=None,
                 density_bias: float = 0.5, roughness_bias: float = -1.0) -> None:
        super().__init__()
        if (base_mlp_num_layers, base_mlp_layer_width, tuple(skip_connections), head_mlp_num_layers,
                head_mlp_layer_width) != (8, 256, (4,), 1, 128):
            raise ValueError("the fused sm_100a field kernels implement the 8x256 (skip 4) + 1x128 architecture "
                             "that reflect_sampling_nerf_model.py:98-106 instantiates")
        if spatial_distortion is not None:
            raise ValueError("spatial_distortion is always None in the reference (model.py:103-106)")
        if density_bias != 0.5:
            raise ValueError("density_bias is compiled into the kernels as 0.5 (field.py:46)")
        w, e = self.WIDTH, self.ENC_DIM
        # construction order = field.py:54-86, so a seeded init reproduces the reference's parameters
        self.mlp_base = _Layers([(e, w)] + [(w + e if i == 4 else w, w) for i in range(1, 8)])
        self.field_output_density = _Head(w, 1)
        self.density_bias = density_bias
        self.field_output_low = _Head(w, 3)          # present, never used
        self.field_output_bottleneck = _Head(w, w)
        self.mlp_mid = _Layers([(self.IDE_DIM + w, self.MID_WIDTH)])
        self.field_output_mid = _Head(self.MID_WIDTH, 3)
        self.field_output_normals = _Head(w, 3)
        self.field_output_roughness = _Head(w, 1)
        self.roughness_bias = roughness_bias         # stored, never used (App. B Q2)
        self.field_output_diff = _Head(w, 3)
        self.field_output_tint = _Head(w, 3)
        self._packed: Optional[Tuple[Tensor, Tensor]] = None
        self._packed_key = None
        self._packed_t: Optional[Tuple[Tensor, Tensor]] = None
        # training state (see train_path.py): gradient blob of the wgrad kernel for the backward in flight,
        # the reusable dY stash, and the data-parallel world size for the flat gradient all-reduce
        self._grad_blob: Optional[Tensor] = None
        self._dy_buffer: Optional[Tensor] = None
        self.dp_world_size = 1

    # ------------------------------------------------------------------------------------ derived state
    def _version_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _repack(self) -> None:
        key = self._version_key()
        if self._packed is None or key != self._packed_key:
            with torch.no_grad():
                wblob, bias, wblob_t, wd = ops.pack_field(dict(self.named_parameters()))   # one kernel (csrc/pack.cu)
            self._packed, self._packed_t, self._packed_key = (wblob, bias), (wblob_t, wd), key

    def packed(self) -> Tuple[Tensor, Tensor]:
        """(bf16 operand blob, fp32 bias vector) for the current parameter values."""
        self._repack()
        return self._packed

    def packed_t(self) -> Tuple[Tensor, Tensor]:
        """(transposed bf16 operand blob of the dgrad chains, bf16 density-head row)."""
        self._repack()
        return self._packed_t

    # ------------------------------------------------------------------------------------ evaluation
    def evaluate_samples(self, origins: Tensor, directions: Tensor, pixel_area: Tensor, bins: Tensor
                         ) -> Dict[str, Tensor]:
        """One fused pass over every frustum sample of a ray batch (field.py:90-186 + components.py:52-140).
        origins/directions [N,3], pixel_area [N,1], bins [N,S+1] -> per-sample tensors [N,S,*]."""
        wblob, bias = self.packed()
        sigma, feat = ops.field_forward(wblob, bias, origins, directions, pixel_area, bins)
        return {"density": sigma, "feat": feat}

    def get_inf_color(self, directions: Tensor, sqradius: Tensor) -> Tensor:
        """field.py:190-201."""
        wblob, bias = self.packed()
        return ops.field_inf_color(wblob, bias, directions, sqradius)
