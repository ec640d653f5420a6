"""Random-init field parameters with the reference's module names and PyTorch's default nn.Linear init
(kaiming-uniform a=sqrt(5)), without needing nerfstudio or the oracle: used by benchmarks and by the
drop-in field when it is constructed stand-alone.  Shapes: reflect_sampling_nerf_field.py:54-86."""
from __future__ import annotations

from collections import OrderedDict

import torch
from torch import nn

FIELD_LINEARS = OrderedDict([
    *[(f"mlp_base.layers.{i}", (256, 99 if i == 0 else 355 if i == 4 else 256)) for i in range(8)],
    ("field_output_density.net", (1, 256)),
    ("field_output_low.net", (3, 256)),
    ("field_output_bottleneck.net", (256, 256)),
    ("mlp_mid.layers.0", (128, 290)),
    ("field_output_mid.net", (3, 128)),
    ("field_output_normals.net", (3, 256)),
    ("field_output_roughness.net", (1, 256)),
    ("field_output_diff.net", (3, 256)),
    ("field_output_tint.net", (3, 256)),
])


def random_field_state(device="cpu"):
    sd = OrderedDict()
    for name, (out_f, in_f) in FIELD_LINEARS.items():
        lin = nn.Linear(in_f, out_f)
        sd[name + ".weight"] = lin.weight.detach().to(device)
        sd[name + ".bias"] = lin.bias.detach().to(device)
    return sd
