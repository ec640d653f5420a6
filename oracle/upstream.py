"""ORACLE (test infrastructure): restatement of the upstream nerfstudio primitives
that the reference's hot path calls.  See oracle/__init__.py for the pinning status.

nerfstudio is NOT vendored in /root/reference (pyproject.toml:6 `nerfstudio >= 0.3.0`,
no lock) and is not installed in this image, so its published 0.3.x / 1.0.x behaviour
is restated here, anchored on the reference's call sites:

  RayBundle / RaySamples / Frustums   model.py:18,283-289  field.py:14,93  components.py:10
  conical_frustum_to_gaussian         field.py:12,93 (via Frustums.get_gaussian_blob)
  NeRFEncoding                        model.py:98-100  field.py:129-131
  MLP, FieldHead & friends            field.py:54-86
  Field.get_normals                   field.py:146-147
  Spaced/Uniform/PDF samplers         model.py:109-112,148,182,292,317  components.py:14-36
  renderers                           model.py:117-124 and every renderer_* call
  Model / ModelConfig / collider      model.py:34,39,78,89-95

Numerics decision "Q-exact" (SURVEY.md §7 item 5, Appendix B): `torch.sum` over the
sample axis in PDFSampler uses ATen's vectorised cascade whose association order
depends on the host's SIMD width, so it cannot be a bit-exact target.  SUM_MODE
selects the variant:  "fp64" (default, the documented oracle variant: accumulate in
double, round once -- what the CUDA kernel reproduces) or "literal" (`torch.sum`).
`torch.cumsum` on CPU already is double-accumulate-then-round, and is kept literal.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, Optional, Tuple, Type, Union

import torch
from torch import Tensor, nn

SUM_MODE = "fp64"  # "fp64" | "literal"


def _row_sum(x: Tensor) -> Tensor:
    if SUM_MODE == "literal":
        return torch.sum(x, dim=-1, keepdim=True)
    return torch.sum(x.double(), dim=-1, keepdim=True).to(x.dtype)


# --------------------------------------------------------------------------- math
@dataclass
class Gaussians:
    mean: Tensor
    cov: Tensor


def compute_3D_gaussian(directions, means, dir_variance, radius_variance) -> Gaussians:
    """Appendix A.1: cov = dir_var * d d^T + rad_var * (I - d (d/|d|^2)^T)."""
    dir_outer = directions[..., :, None] * directions[..., None, :]
    eye = torch.eye(directions.shape[-1], device=directions.device)
    mag_sq = torch.clamp(torch.sum(directions**2, dim=-1, keepdim=True), min=1e-10)
    null_outer = eye - directions[..., :, None] * (directions / mag_sq)[..., None, :]
    cov = dir_variance[..., None] * dir_outer + radius_variance[..., None] * null_outer
    return Gaussians(mean=means, cov=cov)


def conical_frustum_to_gaussian(origins, directions, starts, ends, radius) -> Gaussians:
    """mip-NeRF eq. 7 (stable form).  Appendix A.1."""
    mu = (starts + ends) / 2.0
    hw = (ends - starts) / 2.0
    means = origins + directions * (mu + (2.0 * mu * hw**2.0) / (3.0 * mu**2.0 + hw**2.0))
    dir_variance = (hw**2) / 3 - (4 / 15) * ((hw**4 * (12 * mu**2 - hw**2)) / (3 * mu**2 + hw**2) ** 2)
    radius_variance = radius**2 * ((mu**2) / 4 + (5 / 12) * hw**2 - 4 / 15 * (hw**4) / (3 * mu**2 + hw**2))
    return compute_3D_gaussian(directions, means, dir_variance, radius_variance)


def expected_sin(x_means: Tensor, x_vars: Tensor) -> Tensor:
    return torch.exp(-0.5 * x_vars) * torch.sin(x_means)


def safe_normalize(vectors: Tensor, eps: float = 1e-10) -> Tensor:
    return vectors / (torch.norm(vectors, dim=-1, keepdim=True) + eps)


# --------------------------------------------------------------------------- rays
@dataclass
class Frustums:
    origins: Tensor      # [..., 3]
    directions: Tensor   # [..., 3]
    starts: Tensor       # [..., 1]
    ends: Tensor         # [..., 1]
    pixel_area: Tensor   # [..., 1]
    offsets: Optional[Tensor] = None

    def get_positions(self) -> Tensor:
        pos = self.origins + self.directions * (self.starts + self.ends) / 2
        if self.offsets is not None:
            pos = pos + self.offsets
        return pos

    def get_gaussian_blob(self) -> Gaussians:
        cone_radius = torch.sqrt(self.pixel_area) / 1.7724538509055159
        if self.offsets is not None:
            raise NotImplementedError()
        return conical_frustum_to_gaussian(
            origins=self.origins, directions=self.directions,
            starts=self.starts, ends=self.ends, radius=cone_radius)


@dataclass
class RaySamples:
    frustums: Frustums
    camera_indices: Optional[Tensor] = None
    deltas: Optional[Tensor] = None
    spacing_starts: Optional[Tensor] = None
    spacing_ends: Optional[Tensor] = None
    spacing_to_euclidean_fn: Optional[Callable] = None
    metadata: Optional[Dict[str, Tensor]] = None
    times: Optional[Tensor] = None

    @property
    def shape(self):
        return self.frustums.starts.shape[:-1]

    def get_weights(self, densities: Tensor) -> Tensor:
        """Appendix A.5."""
        delta_density = self.deltas * densities
        alphas = 1 - torch.exp(-delta_density)
        transmittance = torch.cumsum(delta_density[..., :-1, :], dim=-2)
        transmittance = torch.cat(
            [torch.zeros((*transmittance.shape[:1], 1, 1), device=densities.device), transmittance], dim=-2)
        transmittance = torch.exp(-transmittance)
        weights = alphas * transmittance
        return torch.nan_to_num(weights)


@dataclass
class RayBundle:
    origins: Tensor
    directions: Tensor
    pixel_area: Tensor
    camera_indices: Optional[Tensor] = None
    nears: Optional[Tensor] = None
    fars: Optional[Tensor] = None
    metadata: Dict[str, Tensor] = field(default_factory=dict)
    times: Optional[Tensor] = None

    def __len__(self) -> int:
        return int(self.origins.numel() // self.origins.shape[-1])

    @property
    def shape(self):
        return self.origins.shape[:-1]

    def get_ray_samples(self, bin_starts, bin_ends, spacing_starts=None, spacing_ends=None,
                        spacing_to_euclidean_fn=None) -> RaySamples:
        """Appendix A.2: frustum fields are the bundle's [N,1,.] broadcast to [N,S,.]."""
        deltas = bin_ends - bin_starts
        batch = torch.broadcast_shapes(bin_starts.shape[:-1], self.origins[..., None, :].shape[:-1])

        def bc(x: Optional[Tensor]) -> Optional[Tensor]:
            return None if x is None else x.expand(*batch, x.shape[-1])

        frustums = Frustums(
            origins=bc(self.origins[..., None, :]),
            directions=bc(self.directions[..., None, :]),
            starts=bc(bin_starts), ends=bc(bin_ends),
            pixel_area=bc(self.pixel_area[..., None, :]))
        return RaySamples(
            frustums=frustums,
            camera_indices=None if self.camera_indices is None else bc(self.camera_indices[..., None, :]),
            deltas=bc(deltas), spacing_starts=bc(spacing_starts), spacing_ends=bc(spacing_ends),
            spacing_to_euclidean_fn=spacing_to_euclidean_fn, metadata=None,
            times=None if self.times is None else bc(self.times[..., None, :]))


# --------------------------------------------------------------------------- encodings
class FieldComponent(nn.Module):
    def __init__(self, in_dim: Optional[int] = None, out_dim: Optional[int] = None) -> None:
        super().__init__()
        self.in_dim = in_dim
        self.out_dim = out_dim

    def get_out_dim(self) -> int:
        if self.out_dim is None:
            raise ValueError("Output dimension has not been set")
        return self.out_dim


class Encoding(FieldComponent):
    def __init__(self, in_dim: int) -> None:
        if in_dim <= 0:
            raise ValueError("Input dimension should be greater than zero")
        super().__init__(in_dim=in_dim)


class Identity(Encoding):
    def get_out_dim(self) -> int:
        return self.in_dim

    def forward(self, in_tensor):
        return in_tensor


class SHEncoding(Encoding):  # imported by model.py:20, never used
    def __init__(self, levels: int = 4, implementation: str = "torch") -> None:
        super().__init__(in_dim=3)
        self.levels = levels


class NeRFEncoding(Encoding):
    """Appendix A.4 (integrated positional encoding when `covs` is given)."""

    def __init__(self, in_dim, num_frequencies, min_freq_exp, max_freq_exp, include_input=False,
                 implementation="torch") -> None:
        super().__init__(in_dim)
        self.num_frequencies = num_frequencies
        self.min_freq = min_freq_exp
        self.max_freq = max_freq_exp
        self.include_input = include_input

    def get_out_dim(self) -> int:
        out = self.in_dim * self.num_frequencies * 2
        return out + self.in_dim if self.include_input else out

    def forward(self, in_tensor: Tensor, covs: Optional[Tensor] = None) -> Tensor:
        scaled = 2 * torch.pi * in_tensor
        freqs = 2 ** torch.linspace(self.min_freq, self.max_freq, self.num_frequencies, device=in_tensor.device)
        scaled_inputs = scaled[..., None] * freqs
        scaled_inputs = scaled_inputs.view(*scaled_inputs.shape[:-2], -1)
        if covs is None:
            enc = torch.sin(torch.cat([scaled_inputs, scaled_inputs + torch.pi / 2.0], dim=-1))
        else:
            input_var = torch.diagonal(covs, dim1=-2, dim2=-1)[..., :, None] * freqs[None, :] ** 2
            input_var = input_var.reshape((*input_var.shape[:-2], -1))
            enc = expected_sin(
                torch.cat([scaled_inputs, scaled_inputs + torch.pi / 2.0], dim=-1),
                torch.cat(2 * [input_var], dim=-1))
        if self.include_input:
            enc = torch.cat([enc, in_tensor], dim=-1)
        return enc


# --------------------------------------------------------------------------- MLP + heads
class MLP(FieldComponent):
    """Appendix A.7: at a skip layer the ORIGINAL input is concatenated FIRST."""

    def __init__(self, in_dim, num_layers, layer_width, out_dim=None, skip_connections=None,
                 activation=nn.ReLU(), out_activation=None, implementation="torch") -> None:
        super().__init__()
        assert in_dim > 0
        self.in_dim = in_dim
        self.out_dim = out_dim if out_dim is not None else layer_width
        self.num_layers = num_layers
        self.layer_width = layer_width
        self.skip_connections = skip_connections
        self._skip = set(skip_connections) if skip_connections else set()
        self.activation = activation
        self.out_activation = out_activation
        layers = []
        if num_layers == 1:
            layers.append(nn.Linear(in_dim, self.out_dim))
        else:
            for i in range(num_layers - 1):
                if i == 0:
                    assert i not in self._skip
                    layers.append(nn.Linear(in_dim, layer_width))
                elif i in self._skip:
                    layers.append(nn.Linear(layer_width + in_dim, layer_width))
                else:
                    layers.append(nn.Linear(layer_width, layer_width))
            layers.append(nn.Linear(layer_width, self.out_dim))
        self.layers = nn.ModuleList(layers)

    def forward(self, in_tensor: Tensor) -> Tensor:
        x = in_tensor
        for i, layer in enumerate(self.layers):
            if i in self._skip:
                x = torch.cat([in_tensor, x], -1)
            x = layer(x)
            if self.activation is not None and i < len(self.layers) - 1:
                x = self.activation(x)
        if self.out_activation is not None:
            x = self.out_activation(x)
        return x


class FieldHeadNames:
    RGB = "rgb"
    DENSITY = "density"
    NORMALS = "normals"
    PRED_NORMALS = "pred_normals"


class FieldHead(FieldComponent):
    def __init__(self, out_dim, field_head_name, in_dim=None, activation=None) -> None:
        super().__init__()
        self.out_dim = out_dim
        self.activation = activation
        self.field_head_name = field_head_name
        self.net = None
        if in_dim is not None:
            self.in_dim = in_dim
            self.net = nn.Linear(in_dim, out_dim)

    def forward(self, in_tensor: Tensor) -> Tensor:
        out = self.net(in_tensor)
        if self.activation:
            out = self.activation(out)
        return out


class DensityFieldHead(FieldHead):
    def __init__(self, in_dim=None, activation=nn.Softplus()) -> None:
        super().__init__(in_dim=in_dim, out_dim=1, field_head_name=FieldHeadNames.DENSITY, activation=activation)


class RGBFieldHead(FieldHead):
    def __init__(self, in_dim=None, activation=nn.Sigmoid()) -> None:
        super().__init__(in_dim=in_dim, out_dim=3, field_head_name=FieldHeadNames.RGB, activation=activation)


class PredNormalsFieldHead(FieldHead):
    def __init__(self, in_dim=None, activation=nn.Tanh()) -> None:
        super().__init__(in_dim=in_dim, out_dim=3, field_head_name=FieldHeadNames.PRED_NORMALS,
                         activation=activation)

    def forward(self, in_tensor: Tensor) -> Tensor:
        out = super().forward(in_tensor)
        return torch.nn.functional.normalize(out, dim=-1)


class SpatialDistortion(nn.Module):
    pass


class Field(nn.Module):
    """Only what the reference field uses: the density-gradient normals (A.8)."""

    def __init__(self) -> None:
        super().__init__()
        self._sample_locations = None
        self._density_before_activation = None

    def get_normals(self) -> Tensor:
        assert self._sample_locations is not None
        assert self._density_before_activation is not None
        normals = torch.autograd.grad(
            self._density_before_activation, self._sample_locations,
            grad_outputs=torch.ones_like(self._density_before_activation), retain_graph=True)[0]
        return -torch.nn.functional.normalize(normals, dim=-1)


# --------------------------------------------------------------------------- samplers
class Sampler(nn.Module):
    def __init__(self, num_samples: Optional[int] = None) -> None:
        super().__init__()
        self.num_samples = num_samples

    def forward(self, *args, **kwargs):
        return self.generate_ray_samples(*args, **kwargs)


class SpacedSampler(Sampler):
    """Appendix A.2.  `t_rand` may be injected (parity tests) instead of drawn."""

    def __init__(self, spacing_fn, spacing_fn_inv, num_samples=None, train_stratified=True,
                 single_jitter=False) -> None:
        super().__init__(num_samples=num_samples)
        self.train_stratified = train_stratified
        self.single_jitter = single_jitter
        self.spacing_fn = spacing_fn
        self.spacing_fn_inv = spacing_fn_inv
        self.injected_rand: Optional[Tensor] = None

    def generate_ray_samples(self, ray_bundle: RayBundle = None, num_samples: Optional[int] = None) -> RaySamples:
        assert ray_bundle is not None and ray_bundle.nears is not None and ray_bundle.fars is not None
        num_samples = num_samples or self.num_samples
        num_rays = ray_bundle.origins.shape[0]
        bins = torch.linspace(0.0, 1.0, num_samples + 1).to(ray_bundle.origins.device)[None, ...]
        if self.train_stratified and self.training:
            if self.injected_rand is not None:
                t_rand = self.injected_rand
            elif self.single_jitter:
                t_rand = torch.rand((num_rays, 1), dtype=bins.dtype, device=bins.device)
            else:
                t_rand = torch.rand((num_rays, num_samples + 1), dtype=bins.dtype, device=bins.device)
            centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
            upper = torch.cat([centers, bins[..., -1:]], -1)
            lower = torch.cat([bins[..., :1], centers], -1)
            bins = lower + (upper - lower) * t_rand
        s_near, s_far = (self.spacing_fn(x) for x in (ray_bundle.nears, ray_bundle.fars))

        def spacing_to_euclidean_fn(x):
            return self.spacing_fn_inv(x * s_far + (1 - x) * s_near)

        euclid = spacing_to_euclidean_fn(bins)
        return ray_bundle.get_ray_samples(
            bin_starts=euclid[..., :-1, None], bin_ends=euclid[..., 1:, None],
            spacing_starts=bins[..., :-1, None], spacing_ends=bins[..., 1:, None],
            spacing_to_euclidean_fn=spacing_to_euclidean_fn)


class UniformSampler(SpacedSampler):
    def __init__(self, num_samples=None, train_stratified=True, single_jitter=False) -> None:
        super().__init__(num_samples=num_samples, spacing_fn=lambda x: x, spacing_fn_inv=lambda x: x,
                         train_stratified=train_stratified, single_jitter=single_jitter)


class UniformLinDispPiecewiseSampler(SpacedSampler):  # imported by model.py:24, never used
    def __init__(self, num_samples=None, train_stratified=True, single_jitter=False) -> None:
        super().__init__(
            num_samples=num_samples,
            spacing_fn=lambda x: torch.where(x < 1, x / 2, 1 - 1 / (2 * x)),
            spacing_fn_inv=lambda x: torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x)),
            train_stratified=train_stratified, single_jitter=single_jitter)


class PDFSampler(Sampler):
    """Appendix A.3."""

    def __init__(self, num_samples=None, train_stratified=True, single_jitter=False, include_original=True,
                 histogram_padding=0.01) -> None:
        super().__init__(num_samples=num_samples)
        self.train_stratified = train_stratified
        self.include_original = include_original
        self.histogram_padding = histogram_padding
        self.single_jitter = single_jitter
        self.injected_rand: Optional[Tensor] = None
        self.last_inds: Optional[Tensor] = None  # exposed for the bit-exact index test

    def generate_ray_samples(self, ray_bundle: RayBundle = None, ray_samples: RaySamples = None,
                             weights: Tensor = None, num_samples: Optional[int] = None,
                             eps: float = 1e-5) -> RaySamples:
        if ray_samples is None or ray_bundle is None:
            raise ValueError("ray_samples and ray_bundle must be provided")
        assert weights is not None
        num_samples = num_samples or self.num_samples
        num_bins = num_samples + 1
        weights = weights[..., 0] + self.histogram_padding
        weights_sum = _row_sum(weights)
        padding = torch.relu(eps - weights_sum)
        weights = weights + padding / weights.shape[-1]
        weights_sum = weights_sum + padding
        pdf = weights / weights_sum
        cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, dim=-1))
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
        if self.train_stratified and self.training:
            u = torch.linspace(0.0, 1.0 - (1.0 / num_bins), steps=num_bins, device=cdf.device)
            u = u.expand(size=(*cdf.shape[:-1], num_bins))
            if self.injected_rand is not None:
                rand = self.injected_rand / num_bins
            elif self.single_jitter:
                rand = torch.rand((*cdf.shape[:-1], 1), device=cdf.device) / num_bins
            else:
                rand = torch.rand((*cdf.shape[:-1], num_samples + 1), device=cdf.device) / num_bins
            u = u + rand
        else:
            u = torch.linspace(0.0, 1.0 - (1.0 / num_bins), steps=num_bins, device=cdf.device)
            u = u + 1.0 / (2 * num_bins)
            u = u.expand(size=(*cdf.shape[:-1], num_bins))
        u = u.contiguous()
        existing_bins = torch.cat(
            [ray_samples.spacing_starts[..., 0], ray_samples.spacing_ends[..., -1:, 0]], dim=-1)
        inds = torch.searchsorted(cdf, u, side="right")
        self.last_inds = inds
        below = torch.clamp(inds - 1, 0, existing_bins.shape[-1] - 1)
        above = torch.clamp(inds, 0, existing_bins.shape[-1] - 1)
        cdf_g0 = torch.gather(cdf, -1, below)
        bins_g0 = torch.gather(existing_bins, -1, below)
        cdf_g1 = torch.gather(cdf, -1, above)
        bins_g1 = torch.gather(existing_bins, -1, above)
        t = torch.clip(torch.nan_to_num((u - cdf_g0) / (cdf_g1 - cdf_g0), 0), 0, 1)
        bins = bins_g0 + t * (bins_g1 - bins_g0)
        if self.include_original:
            bins, _ = torch.sort(torch.cat([existing_bins, bins], -1), -1)
        bins = bins.detach()
        euclid = ray_samples.spacing_to_euclidean_fn(bins)
        return ray_bundle.get_ray_samples(
            bin_starts=euclid[..., :-1, None], bin_ends=euclid[..., 1:, None],
            spacing_starts=bins[..., :-1, None], spacing_ends=bins[..., 1:, None],
            spacing_to_euclidean_fn=ray_samples.spacing_to_euclidean_fn)


# --------------------------------------------------------------------------- renderers
WHITE = torch.tensor([1.0, 1.0, 1.0])
BLACK = torch.tensor([0.0, 0.0, 0.0])


class RGBRenderer(nn.Module):
    """Appendix A.6."""

    def __init__(self, background_color: Union[str, Tensor] = "random") -> None:
        super().__init__()
        self.background_color = background_color

    @classmethod
    def get_background_color(cls, background_color, shape, device):
        assert isinstance(background_color, Tensor)
        return background_color.expand(shape).to(device)

    @classmethod
    def combine_rgb(cls, rgb, weights, background_color="random"):
        comp_rgb = torch.sum(weights * rgb, dim=-2)
        accumulated_weight = torch.sum(weights, dim=-2)
        if isinstance(background_color, str) and background_color == "random":
            return comp_rgb
        if isinstance(background_color, str):
            raise NotImplementedError(background_color)
        bg = cls.get_background_color(background_color, shape=comp_rgb.shape, device=comp_rgb.device)
        return comp_rgb + bg * (1.0 - accumulated_weight)

    def blend_background(self, image: Tensor, background_color=None) -> Tensor:
        if image.size(-1) < 4:
            return image
        rgb, opacity = image[..., :3], image[..., 3:]
        if background_color is None:
            background_color = self.background_color
            if isinstance(background_color, str):
                background_color = BLACK.to(rgb.device)
        bg = self.get_background_color(background_color, shape=rgb.shape, device=rgb.device)
        return rgb * opacity + bg.to(rgb.device) * (1 - opacity)

    def blend_background_for_loss_computation(self, pred_image, pred_accumulation, gt_image):
        if isinstance(self.background_color, str):
            raise NotImplementedError("only tensor backgrounds are on the reference path (model.py:117-118)")
        bg = self.get_background_color(self.background_color, shape=pred_image.shape, device=pred_image.device)
        gt_image = self.blend_background(gt_image, background_color=bg)
        return pred_image, gt_image

    def forward(self, rgb, weights, ray_indices=None, num_rays=None, background_color=None):
        if background_color is None:
            background_color = self.background_color
        if not self.training:
            rgb = torch.nan_to_num(rgb)
        rgb = self.combine_rgb(rgb, weights, background_color=background_color)
        if not self.training:
            torch.clamp_(rgb, min=0.0, max=1.0)
        return rgb


class AccumulationRenderer(nn.Module):
    @classmethod
    def forward(cls, weights, ray_indices=None, num_rays=None):
        return torch.sum(weights, dim=-2)


class DepthRenderer(nn.Module):
    def __init__(self, method: str = "median") -> None:
        super().__init__()
        self.method = method

    def forward(self, weights, ray_samples: RaySamples, ray_indices=None, num_rays=None):
        steps = (ray_samples.frustums.starts + ray_samples.frustums.ends) / 2
        if self.method == "median":
            cumulative = torch.cumsum(weights[..., 0], dim=-1)
            split = torch.ones((*weights.shape[:-2], 1), device=weights.device) * 0.5
            idx = torch.searchsorted(cumulative, split, side="left")
            idx = torch.clamp(idx, 0, steps.shape[-2] - 1)
            return torch.gather(steps[..., 0], dim=-1, index=idx)
        if self.method == "expected":
            eps = 1e-10
            depth = torch.sum(weights * steps, dim=-2) / (torch.sum(weights, -2) + eps)
            return torch.clip(depth, steps.min(), steps.max())
        raise NotImplementedError(self.method)


class NormalsRenderer(nn.Module):
    @classmethod
    def forward(cls, normals, weights, normalize: bool = True):
        n = torch.sum(weights * normals, dim=-2)
        if normalize:
            n = safe_normalize(n)
        return n


class SemanticRenderer(nn.Module):
    @classmethod
    def forward(cls, semantics, weights, ray_indices=None, num_rays=None):
        return torch.sum(weights * semantics, dim=-2)


# --------------------------------------------------------------------------- model glue
class MSELoss(nn.MSELoss):
    pass


def scale_dict(dictionary: Dict[Any, Any], coefficients: Dict[str, float]) -> Dict[Any, Any]:
    for key in dictionary:
        if key in coefficients:
            dictionary[key] *= coefficients[key]
    return dictionary


def to_immutable_dict(d: Dict[str, Any]):
    return field(default_factory=lambda: dict(d))


class NearFarCollider(nn.Module):
    """Appendix A.9: eval resets the near plane to 0; a bundle that already carries nears/fars passes through."""

    def __init__(self, near_plane: float, far_plane: float, reset_near_plane: bool = True, **kwargs) -> None:
        super().__init__()
        self.near_plane = near_plane
        self.far_plane = far_plane
        self.reset_near_plane = reset_near_plane

    def set_nears_and_fars(self, ray_bundle: RayBundle) -> RayBundle:
        ones = torch.ones_like(ray_bundle.origins[..., 0:1])
        near_plane = self.near_plane if (self.training or not self.reset_near_plane) else 0
        ray_bundle.nears = ones * near_plane
        ray_bundle.fars = ones * self.far_plane
        return ray_bundle

    def forward(self, ray_bundle: RayBundle) -> RayBundle:
        if ray_bundle.nears is not None and ray_bundle.fars is not None:
            return ray_bundle
        return self.set_nears_and_fars(ray_bundle)


@dataclass
class InstantiateConfig:
    _target: Type

    def setup(self, **kwargs) -> Any:
        return self._target(self, **kwargs)


@dataclass
class ModelConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: Model)
    enable_collider: bool = True
    collider_params: Optional[Dict[str, float]] = to_immutable_dict({"near_plane": 2.0, "far_plane": 6.0})
    loss_coefficients: Dict[str, float] = to_immutable_dict({"rgb_loss_coarse": 1.0, "rgb_loss_fine": 1.0})
    eval_num_rays_per_chunk: int = 4096
    prompt: Optional[str] = None


class Model(nn.Module):
    config: ModelConfig

    def __init__(self, config: ModelConfig, scene_box=None, num_train_data: int = 0, **kwargs) -> None:
        super().__init__()
        self.config = config
        self.scene_box = scene_box
        self.render_aabb = None
        self.num_train_data = num_train_data
        self.kwargs = kwargs
        self.collider = None
        self.populate_modules()
        self.callbacks = None
        self.device_indicator_param = nn.Parameter(torch.empty(0))

    @property
    def device(self):
        return self.device_indicator_param.device

    def populate_modules(self):
        if self.config.enable_collider:
            assert self.config.collider_params is not None
            self.collider = NearFarCollider(
                near_plane=self.config.collider_params["near_plane"],
                far_plane=self.config.collider_params["far_plane"])

    def forward(self, ray_bundle: RayBundle):
        if self.collider is not None:
            ray_bundle = self.collider(ray_bundle)
        return self.get_outputs(ray_bundle)
