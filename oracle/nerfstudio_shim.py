"""ORACLE (test infrastructure): stand-in `nerfstudio` / `nerfacc` / `torchmetrics`
module trees, registered in `sys.modules`, so that the UNMODIFIED reference files under
/root/reference/reflect_sampling_nerf can be imported and executed in this container
(tests/golden/make_golden.py).  Every symbol resolves to the restatement in
`oracle.upstream`; only the names the three hot-path files import are provided
(model.py:14-36, field.py:12-25, components.py:7-12).

`install()` refuses to shadow a real nerfstudio installation.
"""
from __future__ import annotations

import enum
import importlib.util
import sys
import types

import torch
from torch import nn

from . import upstream as U


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so sub-modules can hang off it
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], child, m)
    return m


class _TemporalDistortionKind(enum.Enum):
    DNERF = "dnerf"


class _NoOpMetric(nn.Module):
    def __init__(self, *args, **kwargs) -> None:
        super().__init__()

    def forward(self, *args, **kwargs):
        raise RuntimeError("image metrics are outside the hot path (SURVEY.md §2.2)")


def _ssim_stub(*args, **kwargs):
    raise RuntimeError("image metrics are outside the hot path (SURVEY.md §2.2)")


def install() -> None:
    if "nerfstudio" in sys.modules and not getattr(sys.modules["nerfstudio"], "__rsn_shim__", False):
        raise RuntimeError("a real nerfstudio is already imported; refusing to shadow it")
    if "nerfstudio" not in sys.modules and importlib.util.find_spec("nerfstudio") is not None:
        raise RuntimeError("a real nerfstudio is installed; use it instead of the shim")

    _mod("nerfstudio", __rsn_shim__=True)
    _mod("nerfstudio.cameras")
    _mod("nerfstudio.cameras.rays", RayBundle=U.RayBundle, RaySamples=U.RaySamples, Frustums=U.Frustums)
    _mod("nerfstudio.configs")
    _mod("nerfstudio.configs.config_utils", to_immutable_dict=U.to_immutable_dict)
    _mod("nerfstudio.configs.base_config", InstantiateConfig=U.InstantiateConfig)
    _mod("nerfstudio.field_components")
    _mod("nerfstudio.field_components.encodings", Encoding=U.Encoding, Identity=U.Identity,
         NeRFEncoding=U.NeRFEncoding, SHEncoding=U.SHEncoding)
    _mod("nerfstudio.field_components.field_heads", FieldHead=U.FieldHead, FieldHeadNames=U.FieldHeadNames,
         DensityFieldHead=U.DensityFieldHead, RGBFieldHead=U.RGBFieldHead,
         PredNormalsFieldHead=U.PredNormalsFieldHead)
    _mod("nerfstudio.field_components.temporal_distortions", TemporalDistortionKind=_TemporalDistortionKind)
    _mod("nerfstudio.field_components.mlp", MLP=U.MLP)
    _mod("nerfstudio.field_components.spatial_distortions", SpatialDistortion=U.SpatialDistortion)
    _mod("nerfstudio.fields")
    _mod("nerfstudio.fields.base_field", Field=U.Field)
    _mod("nerfstudio.model_components")
    _mod("nerfstudio.model_components.losses", MSELoss=U.MSELoss)
    _mod("nerfstudio.model_components.ray_samplers", Sampler=U.Sampler, SpacedSampler=U.SpacedSampler,
         UniformSampler=U.UniformSampler, PDFSampler=U.PDFSampler,
         UniformLinDispPiecewiseSampler=U.UniformLinDispPiecewiseSampler)
    _mod("nerfstudio.model_components.renderers", RGBRenderer=U.RGBRenderer,
         AccumulationRenderer=U.AccumulationRenderer, DepthRenderer=U.DepthRenderer,
         NormalsRenderer=U.NormalsRenderer, SemanticRenderer=U.SemanticRenderer)
    _mod("nerfstudio.models")
    _mod("nerfstudio.models.base_model", Model=U.Model, ModelConfig=U.ModelConfig)
    _mod("nerfstudio.utils")
    _mod("nerfstudio.utils.math", conical_frustum_to_gaussian=U.conical_frustum_to_gaussian,
         Gaussians=U.Gaussians, expected_sin=U.expected_sin, safe_normalize=U.safe_normalize)
    _mod("nerfstudio.utils.colors", WHITE=U.WHITE, BLACK=U.BLACK)
    _mod("nerfstudio.utils.misc", scale_dict=U.scale_dict)
    _mod("nerfstudio.utils.colormaps")

    _mod("nerfacc", OccGridEstimator=type("OccGridEstimator", (), {}))  # components.py:7, unused

    _mod("torchmetrics")
    _mod("torchmetrics.functional", structural_similarity_index_measure=_ssim_stub)
    _mod("torchmetrics.image", PeakSignalNoiseRatio=_NoOpMetric)
    _mod("torchmetrics.image.lpip", LearnedPerceptualImagePatchSimilarity=_NoOpMetric)
