"""Data side of the hot path -- SURVEY.md §8 rows f1 (GPU ray generation + pixel sampling) and f4 (Blender-format scene
I/O + the procedurally generated shiny-sphere scene BASELINE.json configs[1..2] are quoted on).

  Cameras            perspective cameras in nerfstudio's convention (camera_to_worlds [V,3,4], OpenGL axes: -z forward,
                     +y up; fx, fy, cx, cy) -- the subset of nerfstudio.cameras.cameras.Cameras the reference's
                     datamanager uses (reflect_sampling_nerf_datamanager.py:49-58 -> RayGenerator -> generate_rays)
  load_blender       transforms_{split}.json + RGBA PNGs -> Cameras + uint8 images (what BlenderDataParserConfig, named at
                     reflect_sampling_nerf_config.py:18,36, parses: focal = W / 2 / tan(camera_angle_x / 2), principal point
                     at the image centre, alpha blended onto white at sampling time)
  write_shiny_sphere the synthetic scene: a unit mirror-like sphere under an analytic environment, cameras on the
                     radius-4 sphere, 100 views -- generator is new code (nothing in the reference)
  RayDataManager     next_train(step) -> (RayBundle, batch): pixel sampling, ray generation and the target-pixel gather are
                     ONE kernel launch on resident uint8 images (csrc/raygen.cu) instead of the CPU dataloader ->
                     PixelSampler -> RayGenerator hop
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import ops
from .rays import RayBundle


@dataclass
class Cameras:
    camera_to_worlds: Tensor      # [V,3,4]
    fx: Tensor                    # [V]
    fy: Tensor
    cx: Tensor
    cy: Tensor
    height: int
    width: int

    def __len__(self) -> int:
        return self.camera_to_worlds.shape[0]

    def to(self, device) -> "Cameras":
        return Cameras(self.camera_to_worlds.to(device), self.fx.to(device), self.fy.to(device), self.cx.to(device),
                       self.cy.to(device), self.height, self.width)

    def intrinsics(self) -> Tensor:
        """[V,4] = fx, fy, cx, cy (the layout rsn_raygen reads)."""
        return torch.stack([self.fx, self.fy, self.cx, self.cy], dim=-1).float().contiguous()

    def generate_rays(self, pixels: Tensor) -> RayBundle:
        """Cameras.generate_rays for explicit (camera, y, x) pixel indices [N,3] (int64) -- csrc/raygen.cu."""
        o, d, area, _, _ = ops.raygen(self.camera_to_worlds, self.intrinsics(), self.height, self.width, pixels.shape[0],
                                      pixels=pixels)
        return RayBundle(origins=o, directions=d, pixel_area=area, camera_indices=pixels[:, 0:1])

    def camera_ray_bundle(self, camera_index: int) -> RayBundle:
        """Every pixel of one camera, [H,W,*] (upstream Cameras.generate_rays(camera_indices=i, keep_shape=True))."""
        h, w = self.height, self.width
        dev = self.camera_to_worlds.device
        yy, xx = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
        pix = torch.stack([torch.full_like(yy, camera_index), yy, xx], dim=-1).reshape(-1, 3)
        b = self.generate_rays(pix)
        return RayBundle(origins=b.origins.view(h, w, 3), directions=b.directions.view(h, w, 3),
                         pixel_area=b.pixel_area.view(h, w, 1), camera_indices=pix[:, 0:1].view(h, w, 1))


def orbit_cameras(n_views: int, radius: float, height: int, width: int, camera_angle_x: float, seed: int = 0,
                  upper_hemisphere: bool = True) -> Cameras:
    """Cameras on a sphere of `radius` looking at the origin (Blender / OpenGL convention), quasi-uniform directions."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(n_views, generator=g)
    z = u * (0.9 if upper_hemisphere else 1.8) - (0.0 if upper_hemisphere else 0.9)      # elevation: sin in [0, .9] / [-.9, .9]
    phi = torch.rand(n_views, generator=g) * (2 * math.pi)
    r = torch.sqrt(1 - z * z)
    pos = radius * torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], dim=-1)       # z-up world, as the Blender scenes
    back = torch.nn.functional.normalize(pos, dim=-1)                                      # camera +z = away from the target
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(back)
    right = torch.nn.functional.normalize(torch.cross(up, back, dim=-1), dim=-1)
    true_up = torch.cross(back, right, dim=-1)
    c2w = torch.stack([right, true_up, back, pos], dim=-1)                                 # columns: x, y, z axes, position
    focal = 0.5 * width / math.tan(0.5 * camera_angle_x)
    one = torch.ones(n_views)
    return Cameras(c2w.float(), one * focal, one * focal, one * (width / 2.0), one * (height / 2.0), height, width)


# ------------------------------------------------------------------------------------------ Blender format
def load_blender(root: str, split: str = "train", device="cpu") -> Tuple[Cameras, Tensor]:
    """-> (Cameras, images uint8 [V,H,W,4 or 3]).  Layout: <root>/transforms_<split>.json with `camera_angle_x` and
    `frames[].file_path` (no extension) / `frames[].transform_matrix` (4x4 camera-to-world)."""
    from PIL import Image
    with open(os.path.join(root, f"transforms_{split}.json")) as f:
        meta = json.load(f)
    poses, images = [], []
    for frame in meta["frames"]:
        fname = os.path.join(root, frame["file_path"].replace("./", "") + ".png")
        images.append(np.array(Image.open(fname), dtype=np.uint8))
        poses.append(np.array(frame["transform_matrix"], dtype=np.float32))
    imgs = torch.from_numpy(np.stack(images))
    h, w = imgs.shape[1:3]
    c2w = torch.from_numpy(np.stack(poses))[:, :3].contiguous()
    focal = 0.5 * w / math.tan(0.5 * float(meta["camera_angle_x"]))
    one = torch.ones(len(poses))
    cams = Cameras(c2w, one * focal, one * focal, one * (w / 2.0), one * (h / 2.0), h, w)
    return cams.to(device), imgs.to(device)


def _environment(r: Tensor) -> Tensor:
    """Analytic environment radiance for direction r [...,3] (z up): sky gradient, a sun lobe and a checker floor."""
    z = r[..., 2:3]
    sky = torch.cat([0.35 + 0.25 * z, 0.55 + 0.25 * z, 0.85 + 0.15 * z], dim=-1)
    sun_dir = torch.nn.functional.normalize(torch.tensor([0.6, 0.3, 0.74], device=r.device), dim=0)
    sun = torch.clamp((r * sun_dir).sum(-1, keepdim=True), min=0.0) ** 64
    t = 1.0 / torch.clamp(-z, min=1e-3)                               # floor plane z = -1 seen along r
    fx, fy = r[..., 0:1] * t, r[..., 1:2] * t
    checker = ((torch.floor(fx * 2) + torch.floor(fy * 2)) % 2)
    floor = 0.25 + 0.5 * checker * torch.tensor([0.9, 0.7, 0.5], device=r.device)
    env = torch.where(z >= 0, sky, floor * torch.exp(-0.05 * t))
    return torch.clamp(env + sun * torch.tensor([1.0, 0.95, 0.8], device=r.device), 0.0, 1.0)


@torch.no_grad()
def render_shiny_sphere(cams: Cameras, index: int, device="cpu") -> Tensor:
    """Ground-truth RGBA uint8 [H,W,4] of camera `index`: analytic ray / unit-sphere intersection, colour = 0.25 diffuse
    base + 0.75 mirror reflection of the environment, alpha = hit."""
    h, w = cams.height, cams.width
    c2w = cams.camera_to_worlds[index].to(device).double()
    yy, xx = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")
    x = (xx.double() + 0.5 - float(cams.cx[index])) / float(cams.fx[index])
    y = -(yy.double() + 0.5 - float(cams.cy[index])) / float(cams.fy[index])
    dcam = torch.stack([x, y, -torch.ones_like(x)], dim=-1)
    d = torch.nn.functional.normalize(dcam @ c2w[:, :3].T, dim=-1)
    o = c2w[:, 3]
    b = (d * o).sum(-1)
    disc = b * b - ((o * o).sum() - 1.0)
    hit = disc > 0
    t = -b - torch.sqrt(torch.clamp(disc, min=0.0))
    p = o + t[..., None] * d
    n = torch.nn.functional.normalize(p, dim=-1)
    refl = d - 2.0 * (d * n).sum(-1, keepdim=True) * n
    base = torch.tensor([0.8, 0.2, 0.2], device=device, dtype=torch.float64)
    lambert = torch.clamp((n * torch.tensor([0.6, 0.3, 0.74], device=device, dtype=torch.float64)).sum(-1, keepdim=True), min=0.1)
    rgb = 0.25 * base * lambert + 0.75 * _environment(refl.float()).double()
    rgba = torch.cat([torch.where(hit[..., None], rgb, torch.zeros_like(rgb)), hit[..., None].double()], dim=-1)
    return torch.round(torch.clamp(rgba, 0, 1) * 255).to(torch.uint8)


def write_shiny_sphere(root: str, n_views: int = 100, resolution: int = 400, splits=("train", "val", "test"),
                       device="cpu") -> str:
    """Writes the Blender-format synthetic scene (transforms_{split}.json + <split>/r_<i>.png) and returns `root`."""
    from PIL import Image
    angle = 0.6911112070083618                                # the Blender synthetic scenes' camera_angle_x (f = 555.6 px at 400)
    os.makedirs(root, exist_ok=True)
    for si, split in enumerate(splits):
        nv = n_views if split == "train" else max(1, n_views // 10)
        cams = orbit_cameras(nv, 4.0, resolution, resolution, angle, seed=100 + si)
        os.makedirs(os.path.join(root, split), exist_ok=True)
        frames = []
        for i in range(nv):
            img = render_shiny_sphere(cams, i, device).cpu().numpy()
            Image.fromarray(img, mode="RGBA").save(os.path.join(root, split, f"r_{i}.png"))
            m = torch.cat([cams.camera_to_worlds[i], torch.tensor([[0.0, 0.0, 0.0, 1.0]])]).tolist()
            frames.append({"file_path": f"./{split}/r_{i}", "rotation": 0.0, "transform_matrix": m})
        with open(os.path.join(root, f"transforms_{split}.json"), "w") as f:
            json.dump({"camera_angle_x": angle, "frames": frames}, f)
    return root


def shiny_sphere_in_memory(n_views: int = 100, resolution: int = 400, device="cuda", seed: int = 100
                           ) -> Tuple[Cameras, Tensor]:
    """The same scene without touching the disk (bench.py): (Cameras, images uint8 [V,H,W,4]) on `device`."""
    cams = orbit_cameras(n_views, 4.0, resolution, resolution, 0.6911112070083618, seed=seed)
    imgs = torch.stack([render_shiny_sphere(cams, i, device) for i in range(n_views)])
    return cams.to(device), imgs


# ------------------------------------------------------------------------------------------ datamanager
class RayDataManager:
    """next_train / next_eval of the reference's datamanager (reflect_sampling_nerf_datamanager.py:49-58) with the images
    and cameras resident on the GPU: one launch draws the pixels, generates the rays and gathers + alpha-blends the
    targets.  Collider planes are left to the model (nears / fars None), as upstream's RayGenerator does."""

    def __init__(self, cameras: Cameras, images: Tensor, rays_per_batch: int, seed: Optional[int] = None) -> None:
        if images.dtype != torch.uint8 or images.shape[:3] != (len(cameras), cameras.height, cameras.width):
            raise ValueError("images must be uint8 [V,H,W,C] matching the cameras")
        self.cameras, self.images, self.rays_per_batch = cameras, images.contiguous(), rays_per_batch
        self.c2w, self.intr = cameras.camera_to_worlds.float().contiguous(), cameras.intrinsics()
        self.generator = None
        if seed is not None:
            self.generator = torch.Generator(device=images.device)
            self.generator.manual_seed(seed)
        self.train_count = 0

    def next_train(self, step: int = 0) -> Tuple[RayBundle, Dict[str, Tensor]]:
        self.train_count += 1
        n, dev = self.rays_per_batch, self.images.device
        rand = torch.rand(n, 3, device=dev, generator=self.generator)
        o, d, area, pix, target = ops.raygen(self.c2w, self.intr, self.cameras.height, self.cameras.width, n, rand=rand,
                                             images=self.images)
        return RayBundle(origins=o, directions=d, pixel_area=area, camera_indices=pix[:, 0:1]), {"image": target, "indices": pix}

    def eval_image(self, index: int) -> Tuple[RayBundle, Dict[str, Tensor]]:
        """One full frame: ([H,W,*] RayBundle, {"image": float RGBA or RGB [H,W,C] in [0,1]})."""
        return self.cameras.camera_ray_bundle(index), {"image": self.images[index].float() / 255.0}
