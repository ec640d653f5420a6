"""RayBundle / RaySamples / Frustums containers with nerfstudio's field names and the methods the reference calls on
them (`nerfstudio.cameras.rays`; reflect_sampling_nerf_model.py:18,154,188,283-289,296,322;
reflect_sampling_nerf_field.py:93).  When nerfstudio is importable its own classes are re-exported instead, so the
drop-in model accepts upstream bundles; the model only reads attributes (origins, directions, pixel_area, nears, fars),
so either type works.

RaySamples produced by the samplers of `components.py` also carry the [N,S+1] bin arrays they were cut from
(`_rsn_bins`), which is what the kernels consume; `get_weights` and `Frustums.get_gaussian_blob` run on the kernels.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Optional

from torch import Tensor

try:  # pragma: no cover - nerfstudio is not installed in the build image
    from nerfstudio.cameras.rays import Frustums, RayBundle, RaySamples  # type: ignore # noqa: F401
    HAVE_NERFSTUDIO = True
except Exception:  # noqa: BLE001
    HAVE_NERFSTUDIO = False

    @dataclass
    class Gaussians:
        """nerfstudio.utils.math.Gaussians."""
        mean: Tensor
        cov: Tensor

    @dataclass
    class Frustums:
        origins: Tensor
        directions: Tensor
        starts: Tensor
        ends: Tensor
        pixel_area: Tensor
        offsets: Optional[Tensor] = None

        def get_gaussian_blob(self) -> Gaussians:
            """conical_frustum_to_gaussian of every sample: mean [...,3], cov [...,3,3]."""
            from . import ops
            shape = self.starts.shape[:-1]                      # [N, S]
            n, s = shape[0], shape[-1] if len(shape) > 1 else 1
            o = self.origins.reshape(n, s, 3)[:, 0]
            d = self.directions.reshape(n, s, 3)[:, 0]
            a = self.pixel_area.reshape(n, s)[:, 0]
            bins = self.starts.new_empty(n, s + 1)
            bins[:, :s] = self.starts.reshape(n, s)
            bins[:, s] = self.ends.reshape(n, s)[:, -1]
            mean, cov = ops.frustum_gaussians(o, d, a, bins)
            return Gaussians(mean=mean.reshape(*shape, 3), cov=cov.reshape(*shape, 3, 3))

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

        def get_weights(self, densities: Tensor) -> Tensor:
            """alpha compositing weights [N,S,1] -- rsn_composite_fwd / bwd."""
            from . import ops
            bins = getattr(self, "_rsn_bins", None)
            n, s = densities.shape[0], densities.shape[1]
            if bins is not None:
                eu = bins[1]
            else:
                eu = densities.new_empty(n, s + 1)
                eu[:, :s] = self.frustums.starts.reshape(n, s)
                eu[:, s] = self.frustums.ends.reshape(n, s)[:, -1]
            w, _, _, _ = ops.composite(densities.reshape(n, s), eu, None)
            return w[..., None]

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

        def get_ray_samples(self, bin_starts: Tensor, bin_ends: Tensor, spacing_starts: Optional[Tensor] = None,
                            spacing_ends: Optional[Tensor] = None,
                            spacing_to_euclidean_fn: Optional[Callable] = None) -> RaySamples:
            """nerfstudio RayBundle.get_ray_samples: frustum origins / directions / pixel_area are the bundle's
            [N,1,.] views broadcast along the samples."""
            s = bin_starts.shape[-2]
            n = self.origins.shape[0]
            frustums = Frustums(origins=self.origins[:, None, :].expand(n, s, 3),
                                directions=self.directions[:, None, :].expand(n, s, 3), starts=bin_starts, ends=bin_ends,
                                pixel_area=self.pixel_area.reshape(n, 1, 1).expand(n, s, 1))
            return RaySamples(frustums=frustums, camera_indices=self.camera_indices, deltas=bin_ends - bin_starts,
                              spacing_starts=spacing_starts, spacing_ends=spacing_ends,
                              spacing_to_euclidean_fn=spacing_to_euclidean_fn, metadata=self.metadata, times=self.times)
