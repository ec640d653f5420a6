"""Component classes of the drop-in, with the reference's names, constructor arguments and call forms:

  * reflect_sampling_nerf_components.py:14-36   ReciprocalSampler(SpacedSampler)
  * reflect_sampling_nerf_components.py:38-140  IntegratedSHEncoding(Encoding)
  * the upstream classes reflect_sampling_nerf_model.py:98-124 instantiates (restated from nerfstudio 0.3.x/1.0.x,
    SURVEY.md App. A.2-A.6): UniformSampler, PDFSampler, NeRFEncoding, RGBRenderer, AccumulationRenderer, DepthRenderer,
    NormalsRenderer, SemanticRenderer.

Call forms kept: `sampler(ray_bundle)`, `pdf(ray_bundle, ray_samples, weights)`, `renderer_rgb(rgb, weights,
background_color=...)`, `renderer_depth(weights, ray_samples)`, `encoding(x, covs=...)`, `direction_encoding(dirs, rough)`.
Every forward runs a kernel of librsn_b200.so (csrc/sampling.cu, composite.cu, encode.cu); there is no CPU path.  The
model's get_outputs does not go through the renderer / encoding modules -- its passes use the fused kernels
(train_path.py) -- they exist so that code written against the reference's attributes keeps working.
"""
from __future__ import annotations

from typing import Callable, Optional, Union

import torch
from torch import Tensor, nn

from . import ops
from .rays import RayBundle, RaySamples


# ------------------------------------------------------------------------------------------------ samplers
class SpacedSampler(nn.Module):
    """nerfstudio SpacedSampler: stratified bins in spacing space between near and far."""

    _kind, _tan = ops.UNIFORM, 1.0

    def __init__(self, spacing_fn: Optional[Callable] = None, spacing_fn_inv: Optional[Callable] = None,
                 num_samples: Optional[int] = None, train_stratified: bool = True, single_jitter: bool = False) -> None:
        super().__init__()
        self.num_samples = num_samples
        self.train_stratified = train_stratified
        self.single_jitter = single_jitter
        self.spacing_fn = spacing_fn if spacing_fn is not None else (lambda x: x)
        self.spacing_fn_inv = spacing_fn_inv if spacing_fn_inv is not None else (lambda x: x)
        self.injected_rand: Optional[Tensor] = None      # parity tests supply the stratification noise

    # -- what the fused model path uses -----------------------------------------------------------------------------
    @property
    def kind(self) -> int:
        return self._kind

    def noise(self, n: int, device) -> Optional[Tensor]:
        """[n, S+1] (or [n,1]) uniform noise in training mode, None in eval.  An injected tensor with fewer rows (the
        number of bouncing rays a test knows) is zero-padded to the capacity n."""
        if not (self.train_stratified and self.training):
            return None
        if self.injected_rand is not None:
            r = self.injected_rand.to(device)
            if r.shape[0] < n:
                r = torch.cat([r, r.new_zeros(n - r.shape[0], r.shape[1])])
            return r[:n]
        return torch.rand(n, 1 if self.single_jitter else self.num_samples + 1, device=device)

    # -- the upstream call form -------------------------------------------------------------------------------------
    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, num_samples: Optional[int] = None) -> RaySamples:
        assert ray_bundle is not None
        assert ray_bundle.nears is not None and ray_bundle.fars is not None
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        n, dev = ray_bundle.origins.shape[0], ray_bundle.origins.device
        keep, self.num_samples = self.num_samples, num_samples
        try:
            noise = self.noise(n, dev)
        finally:
            self.num_samples = keep
        spacing, euclid = ops.sample_spaced(ray_bundle.nears, ray_bundle.fars, num_samples, self._kind, noise, tan=self._tan)
        return _samples_from_bins(ray_bundle, spacing, euclid, self, self._kind, self._tan)

    def forward(self, *args, **kwargs) -> RaySamples:
        return self.generate_ray_samples(*args, **kwargs)


def _samples_from_bins(ray_bundle, spacing: Tensor, euclid: Tensor, sampler, kind: int, tan: float) -> RaySamples:
    s_near, s_far = sampler.spacing_fn(ray_bundle.nears), sampler.spacing_fn(ray_bundle.fars)
    inv = sampler.spacing_fn_inv

    def spacing_to_euclidean_fn(x):
        return inv(x * s_far + (1 - x) * s_near)

    rs = ray_bundle.get_ray_samples(bin_starts=euclid[..., :-1, None], bin_ends=euclid[..., 1:, None],
                                    spacing_starts=spacing[..., :-1, None], spacing_ends=spacing[..., 1:, None],
                                    spacing_to_euclidean_fn=spacing_to_euclidean_fn)
    # what the kernels consume: the [N,S+1] bin arrays and the spacing description
    object.__setattr__(rs, "_rsn_bins", (spacing, euclid))
    object.__setattr__(rs, "_rsn_spacing", (kind, tan, ray_bundle.nears, ray_bundle.fars, sampler))
    object.__setattr__(rs, "_rsn_bundle", ray_bundle)
    return rs


class UniformSampler(SpacedSampler):
    """nerfstudio UniformSampler: identity spacing (reflect_sampling_nerf_model.py:109)."""

    def __init__(self, num_samples: Optional[int] = None, train_stratified: bool = True, single_jitter: bool = False) -> None:
        super().__init__(spacing_fn=lambda x: x, spacing_fn_inv=lambda x: x, num_samples=num_samples,
                         train_stratified=train_stratified, single_jitter=single_jitter)


class ReciprocalSampler(SpacedSampler):
    """reflect_sampling_nerf_components.py:14-36: s(x) = x / (1/tan + x), s^-1(u) = u / tan / (1 - u)."""

    _kind = ops.RECIPROCAL

    def __init__(self, tan: float = 1.0, num_samples: Optional[int] = None, train_stratified=True, single_jitter=False) -> None:
        super().__init__(spacing_fn=lambda x: x / (1 / tan + x), spacing_fn_inv=lambda x: x / tan / (1 - x),
                         num_samples=num_samples, train_stratified=train_stratified, single_jitter=single_jitter)
        self._tan = float(tan)
        self.tan = tan


class PDFSampler(nn.Module):
    """nerfstudio PDFSampler.  The kernel implements include_original=False, the form the
    reference instantiates (reflect_sampling_nerf_model.py:110,112)."""

    def __init__(self, num_samples: Optional[int] = None, train_stratified: bool = True, single_jitter: bool = False,
                 include_original: bool = True, histogram_padding: float = 0.01) -> None:
        super().__init__()
        if include_original:
            raise NotImplementedError("PDFSampler(include_original=True) is not on the reference's path "
                                      "(reflect_sampling_nerf_model.py:110,112 pass False) and has no kernel")
        if single_jitter:
            raise NotImplementedError("PDFSampler(single_jitter=True) has no kernel")
        self.num_samples = num_samples
        self.train_stratified = train_stratified
        self.single_jitter = single_jitter
        self.include_original = include_original
        self.histogram_padding = histogram_padding
        self.kind = ops.UNIFORM                           # spacing of the samples it refines (set by the model)
        self.injected_rand: Optional[Tensor] = None

    def noise(self, n: int, device) -> Optional[Tensor]:
        if not (self.train_stratified and self.training):
            return None
        if self.injected_rand is not None:
            r = self.injected_rand.to(device)
            if r.shape[0] < n:
                r = torch.cat([r, r.new_zeros(n - r.shape[0], r.shape[1])])
            return r[:n]
        return torch.rand(n, self.num_samples + 1, device=device)

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, ray_samples: Optional[RaySamples] = None,
                             weights: Optional[Tensor] = None, num_bins: Optional[int] = None, eps: float = 1e-5) -> RaySamples:
        if ray_samples is None or ray_bundle is None or weights is None:
            raise ValueError("ray_samples, ray_bundle and weights must be provided")
        if eps != 1e-5:
            raise NotImplementedError("the kernel uses PDFSampler's default eps = 1e-5")
        assert self.num_samples is not None
        n_out = num_bins - 1 if num_bins is not None else self.num_samples     # upstream: num_bins = num_samples + 1
        n, dev = weights.shape[0], weights.device
        bins_in = getattr(ray_samples, "_rsn_bins", None)
        if bins_in is not None:
            spacing_in = bins_in[0]
        else:
            spacing_in = torch.cat([ray_samples.spacing_starts[..., 0], ray_samples.spacing_ends[..., -1:, 0]], dim=-1)
        keep, self.num_samples = self.num_samples, n_out
        try:
            noise = self.noise(n, dev)
        finally:
            self.num_samples = keep
        desc = getattr(ray_samples, "_rsn_spacing", None)
        train = self.train_stratified and self.training
        if desc is not None:
            kind, tan, nears, fars, sampler = desc
            spacing, euclid = ops.pdf_resample(weights, spacing_in, nears, fars, n_out, kind, rand=noise, train=train,
                                               histogram_padding=self.histogram_padding, tan=tan)
            return _samples_from_bins(ray_bundle, spacing, euclid, sampler, kind, tan)
        # foreign RaySamples: resample in spacing space on the kernel, map to Euclidean space with the samples' own function
        zeros, ones = weights.new_zeros(n), weights.new_ones(n)
        spacing, _ = ops.pdf_resample(weights, spacing_in, zeros, ones, n_out, ops.UNIFORM, rand=noise, train=train,
                                      histogram_padding=self.histogram_padding)
        assert ray_samples.spacing_to_euclidean_fn is not None
        euclid = ray_samples.spacing_to_euclidean_fn(spacing)
        rs = ray_bundle.get_ray_samples(bin_starts=euclid[..., :-1, None], bin_ends=euclid[..., 1:, None],
                                        spacing_starts=spacing[..., :-1, None], spacing_ends=spacing[..., 1:, None],
                                        spacing_to_euclidean_fn=ray_samples.spacing_to_euclidean_fn)
        object.__setattr__(rs, "_rsn_bins", (spacing, euclid.contiguous()))
        object.__setattr__(rs, "_rsn_bundle", ray_bundle)
        return rs

    def forward(self, *args, **kwargs) -> RaySamples:
        return self.generate_ray_samples(*args, **kwargs)


# ------------------------------------------------------------------------------------------------ encodings
class Encoding(nn.Module):
    """nerfstudio FieldComponent / Encoding."""

    def __init__(self, in_dim: int) -> None:
        super().__init__()
        if in_dim <= 0:
            raise ValueError("Input dimension should be greater than zero")
        self.in_dim = in_dim

    def get_out_dim(self) -> int:
        raise NotImplementedError


class NeRFEncoding(Encoding):
    """nerfstudio NeRFEncoding in the one configuration the reference builds (reflect_sampling_nerf_model.py:98-100):
    in_dim 3, 16 frequencies 2**linspace(0, 16, 16), include_input; with `covs` = the integrated positional encoding
    (only diag(cov) is read, the variance is NOT scaled by (2 pi)^2)."""

    def __init__(self, in_dim: int = 3, num_frequencies: int = 16, min_freq_exp: float = 0.0, max_freq_exp: float = 16.0,
                 include_input: bool = True, **kwargs) -> None:
        super().__init__(in_dim)
        if (in_dim, num_frequencies, float(min_freq_exp), float(max_freq_exp), include_input) != (3, 16, 0.0, 16.0, True):
            raise ValueError("the kernels implement NeRFEncoding(in_dim=3, num_frequencies=16, min_freq_exp=0, "
                             "max_freq_exp=16, include_input=True) (reflect_sampling_nerf_model.py:98-100)")
        self.num_frequencies, self.min_freq, self.max_freq, self.include_input = 16, 0.0, 16.0, True

    def get_out_dim(self) -> int:
        return 99

    def forward(self, in_tensor: Tensor, covs: Optional[Tensor] = None) -> Tensor:
        return ops.ipe_encode(in_tensor, covs)


class IntegratedSHEncoding(Encoding):
    """reflect_sampling_nerf_components.py:38-140: 34 real-SH-like polynomials of the direction (l = 1, 2, 4, 8, the
    reference's constants verbatim) attenuated by exp(-roughness l (l + 1) / 2)."""

    def __init__(self) -> None:
        super().__init__(in_dim=3)

    def get_out_dim(self) -> int:
        return 34

    @torch.no_grad()
    def pytorch_fwd(self, directions: Tensor, roughness: Tensor) -> Tensor:
        return ops.ide_encode(directions, roughness)

    def forward(self, directions: Tensor, roughness: Tensor) -> Tensor:
        return self.pytorch_fwd(directions, roughness)


# ------------------------------------------------------------------------------------------------ renderers
BackgroundColor = Union[str, Tensor]
WHITE = torch.tensor([1.0, 1.0, 1.0])


def _w2(weights: Tensor) -> Tensor:
    return weights[..., 0] if weights.dim() == 3 else weights


class AccumulationRenderer(nn.Module):
    """sum of the weights along the ray."""

    @classmethod
    def forward(cls, weights: Tensor, ray_indices=None, num_rays=None) -> Tensor:
        acc, _, _ = ops.render_weights(_w2(weights))
        return acc[:, None]


class RGBRenderer(nn.Module):
    """nerfstudio RGBRenderer: sum w rgb (+ background (1 - sum w) unless the background is "random"); eval mode applies
    nan_to_num and clamps to [0, 1]."""

    def __init__(self, background_color: BackgroundColor = "random") -> None:
        super().__init__()
        self.background_color: BackgroundColor = background_color

    def forward(self, rgb: Tensor, weights: Tensor, ray_indices=None, num_rays=None,
                background_color: Optional[BackgroundColor] = None) -> Tensor:
        if background_color is None:
            background_color = self.background_color
        if not self.training:
            rgb = torch.nan_to_num(rgb)
        acc, comp, _ = ops.render_weights(_w2(weights), rgb)
        if not isinstance(background_color, str) or background_color != "random":
            if isinstance(background_color, str):
                background_color = {"white": WHITE, "black": torch.zeros(3)}[background_color]
            comp = comp + background_color.to(comp.device).expand(comp.shape) * (1.0 - acc[:, None])
        if not self.training:
            comp = torch.clamp(comp, min=0.0, max=1.0)
        return comp

    def blend_background(self, image: Tensor, background_color: Optional[BackgroundColor] = None) -> Tensor:
        if image.shape[-1] < 4:
            return image
        rgb, opacity = image[..., :3], image[..., 3:]
        if background_color is None:
            background_color = self.background_color
            if isinstance(background_color, str):
                background_color = WHITE if background_color in ("white", "random", "last_sample") else torch.zeros(3)
        return rgb * opacity + background_color.to(rgb.device) * (1 - opacity)

    def blend_background_for_loss_computation(self, pred_image: Tensor, pred_accumulation: Tensor, gt_image: Tensor):
        """Identity on the prediction for a tensor background; the target is blended only if it has an alpha channel."""
        bg = self.background_color if isinstance(self.background_color, Tensor) else None
        return pred_image, self.blend_background(gt_image, background_color=bg)


class DepthRenderer(nn.Module):
    """nerfstudio DepthRenderer(method="median"): the mid-point of the first sample whose cumulative weight reaches 0.5."""

    def __init__(self, method: str = "median") -> None:
        super().__init__()
        if method != "median":
            raise NotImplementedError("only the default method='median' is on the reference's path")
        self.method = method

    def forward(self, weights: Tensor, ray_samples: RaySamples, ray_indices=None, num_rays=None) -> Tensor:
        bins = getattr(ray_samples, "_rsn_bins", None)
        if bins is not None:
            _, _, depth = ops.render_weights(_w2(weights), bins=bins[1])
        else:
            _, _, depth = ops.render_weights(_w2(weights), starts=ray_samples.frustums.starts, ends=ray_samples.frustums.ends)
        return depth[:, None]


class NormalsRenderer(nn.Module):
    """nerfstudio NormalsRenderer: safe_normalize(sum w n) = v / (|v| + 1e-10)."""

    @classmethod
    def forward(cls, normals: Tensor, weights: Tensor, normalize: bool = True) -> Tensor:
        _, n, _ = ops.render_weights(_w2(weights), normals)
        if normalize:
            n = n / (torch.linalg.norm(n, dim=-1, keepdim=True) + 1e-10)
        return n


class SemanticRenderer(nn.Module):
    """nerfstudio SemanticRenderer: sum w x."""

    @classmethod
    def forward(cls, semantics: Tensor, weights: Tensor, ray_indices=None, num_rays=None) -> Tensor:
        _, out, _ = ops.render_weights(_w2(weights), semantics)
        return out
