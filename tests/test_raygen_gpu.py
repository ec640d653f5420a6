"""Pixel sampling + ray generation + target gather in one launch (csrc/raygen.cu) -- BIT-EXACT against the
oracle restatement of PixelSampler / Cameras.generate_rays / the Blender alpha blend (oracle/cameras.py), the
Blender-format scene round trip and the shiny-sphere generator (§8 f4)."""
import json
import os

import pytest
import torch

from oracle import cameras as C
from reflect_sampling_nerf_b200 import data as D
from reflect_sampling_nerf_b200 import ops

pytestmark = pytest.mark.gpu


def _scene(v=9, res=48):
    cams = D.orbit_cameras(v, 4.0, res, res, 0.6911112070083618, seed=5)
    imgs = torch.stack([D.render_shiny_sphere(cams, i) for i in range(v)])
    return cams, imgs


def test_raygen_bit_exact_vs_oracle():
    cams, imgs = _scene()
    n = 20000
    g = torch.Generator().manual_seed(7)
    rand = torch.rand(n, 3, generator=g)
    pix_ref = C.sample_pixels(rand, len(cams), cams.height, cams.width)
    o_ref, d_ref, a_ref = C.generate_rays(cams.camera_to_worlds, cams.fx, cams.fy, cams.cx, cams.cy, pix_ref)
    t_ref = C.gather_targets(imgs, pix_ref)
    cg = cams.to("cuda")
    o, d, area, pix, target = ops.raygen(cg.camera_to_worlds, cg.intrinsics(), cams.height, cams.width, n, rand=rand.cuda(),
                                         images=imgs.cuda())
    assert torch.equal(pix.cpu(), pix_ref)
    assert torch.equal(o.cpu(), o_ref) and torch.equal(d.cpu(), d_ref) and torch.equal(area.cpu(), a_ref)
    assert torch.equal(target.cpu(), t_ref)
    assert 0.2 < float((t_ref == 1).all(-1).float().mean()) < 0.98          # white background and sphere pixels both occur
    # explicit pixel indices (eval images) take the same arithmetic
    b = cg.generate_rays(pix_ref.cuda())
    assert torch.equal(b.directions.cpu(), d_ref) and torch.equal(b.pixel_area.cpu(), a_ref)
    # and within 1 ulp of the literal upstream tensor expressions
    _, d_up, _ = C.generate_rays_upstream(cams.camera_to_worlds, cams.fx, cams.fy, cams.cx, cams.cy, pix_ref)
    assert float((d.cpu() - d_up).abs().max()) <= 2 ** -22


def test_blender_scene_round_trip_and_datamanager(tmp_path):
    root = D.write_shiny_sphere(str(tmp_path / "sphere"), n_views=6, resolution=40)
    meta = json.load(open(os.path.join(root, "transforms_train.json")))
    assert set(meta) >= {"camera_angle_x", "frames"} and len(meta["frames"]) == 6
    assert set(meta["frames"][0]) >= {"file_path", "transform_matrix"}
    cams, imgs = D.load_blender(root, "train", device="cuda")
    assert imgs.shape == (6, 40, 40, 4) and imgs.dtype == torch.uint8
    ref = D.orbit_cameras(6, 4.0, 40, 40, 0.6911112070083618, seed=100)
    torch.testing.assert_close(cams.camera_to_worlds.cpu(), ref.camera_to_worlds, rtol=0, atol=1e-6)
    assert torch.equal(imgs[2].cpu(), D.render_shiny_sphere(ref, 2))
    dm = D.RayDataManager(cams, imgs, rays_per_batch=4096, seed=3)
    bundle, batch = dm.next_train(0)
    assert bundle.origins.shape == (4096, 3) and batch["image"].shape == (4096, 3) and bundle.nears is None
    torch.testing.assert_close(torch.linalg.norm(bundle.directions, dim=-1), torch.ones(4096, device="cuda"), rtol=0, atol=1e-6)
    # rays aimed at the sphere hit it where the image says so: analytic check of the (camera, pixel) <-> ray pairing
    o, d = bundle.origins, bundle.directions
    bq = (o * d).sum(-1)
    hit = bq * bq - ((o * o).sum(-1) - 1.0) > 0
    is_bg = (batch["image"] == 1.0).all(-1)
    assert float((hit != is_bg).float().mean()) > 0.97
    frame, fb = dm.eval_image(1)
    assert frame.origins.shape == (40, 40, 3) and fb["image"].shape == (40, 40, 4)
