"""ORACLE (test infrastructure -- only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this).

CPU restatement of the data-side hop in front of the hot path, i.e. what
reflect_sampling_nerf_datamanager.py:49-58 calls per training batch:

    batch = self.train_pixel_sampler.sample(image_batch)     # upstream PixelSampler.sample_method
    ray_bundle = self.train_ray_generator(ray_indices)       # upstream RayGenerator -> Cameras.generate_rays

nerfstudio is a third-party dependency absent from /root/reference and from this image (pyproject.toml:6, `>= 0.3.0`, no
lock): PARITY UNPINNED -- restated from nerfstudio 0.3.x / 1.0.x (`cameras/cameras.py::_generate_rays_from_coords`,
perspective cameras without distortion; `cameras/camera_utils.py::normalize_with_norm`;
`data/pixel_samplers.py::PixelSampler.sample_method`; `data/dataparsers/blender_dataparser.py` + `InputDataset.get_image`
for the alpha blend onto white).

Two forms of the ray arithmetic:
  * `generate_rays_upstream`  the literal tensor expressions (torch.sum over the last axis, linalg.vector_norm);
  * `generate_rays`           the same arithmetic with every fp32 operation written out in a fixed order
                              (((a b) + (c d)) + (e f), sqrt, divide) and IEEE-rounded -- the BIT-EXACT target of
                              csrc/raygen.cu.  ATen's reduction order is an implementation detail, and its vectorised CPU
                              sqrt is not correctly rounded (1 ulp off numpy's / CUDA's sqrtf on ~1 % of inputs, measured
                              here), so the square roots are taken in float64 and rounded once (= the IEEE fp32 sqrt).
                              tests/test_oracle_cameras.py measures how far the two forms are apart on CPU (<= 2 ulp).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch
from torch import Tensor

_EPS = float(np.finfo(float).eps * 4.0)          # camera_utils._EPS


def sample_pixels(rand: Tensor, n_views: int, height: int, width: int) -> Tensor:
    """PixelSampler.sample_method: floor(rand[N,3] * (V, H, W)).long() -> (camera, y, x)."""
    return torch.floor(rand * torch.tensor([n_views, height, width], dtype=rand.dtype)).long()


def _coords(pixels: Tensor, fx, fy, cx, cy):
    cam = pixels[:, 0]
    y = pixels[:, 1].float() + 0.5                 # get_image_coords(pixel_offset=0.5)
    x = pixels[:, 2].float() + 0.5
    fx, fy, cx, cy = fx[cam], fy[cam], cx[cam], cy[cam]
    coord = torch.stack([(x - cx) / fx, -(y - cy) / fy], -1)
    coord_x = torch.stack([(x - cx + 1) / fx, -(y - cy) / fy], -1)
    coord_y = torch.stack([(x - cx) / fx, -(y - cy + 1) / fy], -1)
    return cam, torch.stack([coord, coord_x, coord_y], dim=0)          # [3, N, 2]


def generate_rays_upstream(c2w: Tensor, fx, fy, cx, cy, pixels: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Literal form of Cameras._generate_rays_from_coords (perspective): -> origins [N,3], directions [N,3], pixel_area [N,1]."""
    cam, cs = _coords(pixels, fx, fy, cx, cy)
    dirs = torch.stack([cs[..., 0], cs[..., 1], -torch.ones_like(cs[..., 0])], dim=-1)       # [3, N, 3]
    rot = c2w[cam][:, :3, :3]
    dirs = torch.sum(dirs[..., None, :] * rot, dim=-1)
    norm = torch.maximum(torch.linalg.vector_norm(dirs, dim=-1, keepdim=True), torch.tensor([_EPS]).to(dirs))
    dirs = dirs / norm
    origins = c2w[cam][:, :3, 3]
    d = dirs[0]
    dx = torch.sqrt(torch.sum((d - dirs[1]) ** 2, dim=-1))
    dy = torch.sqrt(torch.sum((d - dirs[2]) ** 2, dim=-1))
    return origins, d, (dx * dy)[..., None]


def _sqrt_rn(x: Tensor) -> Tensor:
    """Correctly rounded fp32 square root (float64 sqrt carries 53 >= 2 * 24 + 2 bits: rounding it once is exact)."""
    return torch.sqrt(x.double()).float()


def generate_rays(c2w: Tensor, fx, fy, cx, cy, pixels: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """The same with an explicit fp32 operation order (the kernel's bit-exact target)."""
    cam, cs = _coords(pixels, fx, fy, cx, cy)
    rot = c2w[cam][:, :3, :3]                                           # [N, 3, 3]
    out = []
    for k in range(3):
        u, v = cs[k, :, 0], cs[k, :, 1]
        w = [((u * rot[:, i, 0]) + (v * rot[:, i, 1])) + (-1.0 * rot[:, i, 2]) for i in range(3)]
        nrm = _sqrt_rn(((w[0] * w[0]) + (w[1] * w[1])) + (w[2] * w[2]))
        nrm = torch.maximum(nrm, torch.tensor(_EPS, dtype=nrm.dtype))
        out.append(torch.stack([w[0] / nrm, w[1] / nrm, w[2] / nrm], dim=-1))
    d = out[0]

    def dist(a, b):
        e = a - b
        return _sqrt_rn(((e[:, 0] * e[:, 0]) + (e[:, 1] * e[:, 1])) + (e[:, 2] * e[:, 2]))

    area = dist(d, out[1]) * dist(d, out[2])
    return c2w[cam][:, :3, 3], d, area[..., None]


def gather_targets(images_u8: Tensor, pixels: Tensor) -> Tensor:
    """InputDataset.get_image + PixelSampler gather: uint8 [V,H,W,C] -> float32 [N,3], RGBA blended onto white."""
    px = images_u8[pixels[:, 0], pixels[:, 1], pixels[:, 2]].numpy().astype("float32") / 255.0
    px = torch.from_numpy(px)
    if px.shape[-1] == 4:
        return px[:, :3] * px[:, -1:] + torch.ones(3) * (1.0 - px[:, -1:])
    return px
