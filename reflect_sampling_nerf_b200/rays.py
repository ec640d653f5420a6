"""Minimal RayBundle / RaySamples / Frustums containers with nerfstudio's field names
(`nerfstudio.cameras.rays`, used at reflect_sampling_nerf_model.py:18,283-289).  When nerfstudio is
importable its own classes are re-exported instead, so the drop-in model accepts upstream bundles; the
model only reads attributes (origins, directions, pixel_area, nears, fars), so either type works.
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
    class Frustums:
        origins: Tensor
        directions: Tensor
        starts: Tensor
        ends: Tensor
        pixel_area: Tensor
        offsets: Optional[Tensor] = None

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
