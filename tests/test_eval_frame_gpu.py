"""Eval frame path (BASELINE config C3): chunked full-frame render with the padded
`depth_reflect_fine`, reassembled [H,W,*] outputs, PSNR / SSIM and a working get_image_metrics_and_images
(reflect_sampling_nerf_model.py:432-482 + upstream get_outputs_for_camera_ray_bundle) -- a 64x64 frame against the oracle
rendered in the same chunks."""
import math

import pytest
import torch

from helpers import oracle_model
from oracle import upstream as U
from reflect_sampling_nerf_b200 import data as D
from reflect_sampling_nerf_b200.model import (ReflectSamplingNeRFModel, ReflectSamplingNeRFModelConfig,
                                              structural_similarity_index_measure)

pytestmark = pytest.mark.gpu
SIZES = dict(num_coarse_samples=48, num_importance_samples=48, num_reflect_coarse_samples=24, num_reflect_importance_samples=24)


def test_full_frame_matches_oracle_and_metrics_work():
    res, chunk = 64, 1024
    ref = oracle_model(SIZES, seed=4).eval()
    mine = ReflectSamplingNeRFModel(ReflectSamplingNeRFModelConfig(eval_num_rays_per_chunk=chunk, **SIZES)).cuda().eval()
    mine.field.load_state_dict(ref.field.state_dict(), strict=False)
    cams = D.orbit_cameras(3, 4.0, res, res, 0.6911112070083618, seed=8)
    imgs = torch.stack([D.render_shiny_sphere(cams, i) for i in range(3)])
    dm = D.RayDataManager(cams.to("cuda"), imgs.cuda(), rays_per_batch=chunk)
    frame, batch = dm.eval_image(1)
    out = mine.get_outputs_for_camera_ray_bundle(frame)
    for k in ("mid_rgb_coarse", "mid_rgb_fine", "mid_reflect_coarse", "mid_reflect_fine"):
        assert out[k].shape == (res, res, 3), k
    for k in ("accumulation_fine", "depth_fine", "depth_reflect_fine", "roughness", "mask"):
        assert out[k].shape == (res, res, 1), k
    assert out["weights_fine"].shape == (res, res, 48)
    # oracle: the same chunks through OracleModel.forward (eval mode: collider near plane 0, clamped renderers)
    o, d, a = [t.reshape(-1, t.shape[-1]).cpu() for t in (frame.origins, frame.directions, frame.pixel_area)]
    rows = {k: [] for k in ("mid_rgb_fine", "mid_reflect_fine", "mask", "accumulation_fine")}
    with torch.no_grad():
        for i in range(0, res * res, chunk):
            r = ref(U.RayBundle(origins=o[i:i + chunk], directions=d[i:i + chunk], pixel_area=a[i:i + chunk]))
            for k in rows:
                rows[k].append(r[k])
    refo = {k: torch.cat(v) for k, v in rows.items()}
    agree = refo["mask"] == out["mask"].reshape(-1).cpu()
    assert agree.float().mean() > 0.985
    torch.testing.assert_close(out["mid_rgb_fine"].reshape(-1, 3).cpu(), refo["mid_rgb_fine"], rtol=0, atol=1e-2)
    torch.testing.assert_close(out["mid_reflect_fine"].reshape(-1, 3).cpu()[agree], refo["mid_reflect_fine"][agree], rtol=0, atol=1e-2)
    torch.testing.assert_close(out["accumulation_fine"].reshape(-1, 1).cpu(), refo["accumulation_fine"], rtol=0, atol=1e-2)
    # metrics + images (the reference's version dies on outputs["low_coarse"], App. B Q13)
    metrics, images = mine.get_image_metrics_and_images(out, batch)
    assert set(metrics) >= {"psnr", "coarse_psnr", "fine_psnr", "fine_ssim"}
    gt = batch["image"][..., :3] * batch["image"][..., 3:] + (1 - batch["image"][..., 3:])
    mse = float(torch.mean((gt - torch.clip(out["mid_reflect_fine"], 0, 1)) ** 2))
    assert abs(metrics["fine_psnr"] - 10 * math.log10(1 / mse)) < 1e-3
    assert -1.0 <= metrics["fine_ssim"] <= 1.0
    assert images["img"].shape == (res, 3 * res, 3) and images["accumulation"].shape == (res, 2 * res, 3)
    assert images["depth"].shape == (res, 2 * res, 3)
    x = torch.rand(1, 3, 32, 32, device="cuda")
    assert abs(float(structural_similarity_index_measure(x, x)) - 1.0) < 1e-5
