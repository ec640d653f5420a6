"""Whole-path parity: the drop-in model's get_outputs (CUDA kernels through the C-ABI) against the oracle
model on the same rays, random-init weights and (in train mode) injected stratification noise.
north_star tolerance: rendered RGB within 1e-2 max-abs for the bf16 MLP."""
import pytest
import torch

from helpers import GOLDEN_SIZES, oracle_model, synthetic_rays
from oracle import upstream as U
from reflect_sampling_nerf_b200.model import ReflectSamplingNeRFModel, ReflectSamplingNeRFModelConfig
from reflect_sampling_nerf_b200.rays import RayBundle

pytestmark = pytest.mark.gpu
ATOL_RGB = 1e-2


def _models(sizes, seed=3):
    ref = oracle_model(sizes, seed=seed)
    cfg = ReflectSamplingNeRFModelConfig(**sizes)
    mine = ReflectSamplingNeRFModel(cfg).cuda()
    missing, unexpected = mine.field.load_state_dict(ref.field.state_dict(), strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return ref, mine


def _bundles(n, seed, area):
    o, d, pa, img = synthetic_rays(n, seed, pixel_area=area)
    ref_b = U.RayBundle(origins=o, directions=d, pixel_area=pa)
    my_b = RayBundle(origins=o.cuda(), directions=d.cuda(), pixel_area=pa.cuda())
    return ref_b, my_b, img


@pytest.mark.parametrize("mode", ["eval", "train"])
@pytest.mark.parametrize("sizes", [GOLDEN_SIZES, dict(num_coarse_samples=64, num_importance_samples=64,
                                                       num_reflect_coarse_samples=32,
                                                       num_reflect_importance_samples=32)])
def test_get_outputs_matches_oracle(mode, sizes):
    n = 384
    ref, mine = _models(sizes)
    ref.train(mode == "train")
    mine.train(mode == "train")
    ref_b, my_b, img = _bundles(n, 17, 3.2e-6)
    if mode == "train":
        g = torch.Generator().manual_seed(99)
        jit = dict(uniform=torch.rand(n, sizes["num_coarse_samples"] + 1, generator=g),
                   pdf=torch.rand(n, sizes["num_importance_samples"] + 1, generator=g))
    with torch.set_grad_enabled(mode == "train"):   # the oracle's density-gradient normals need autograd
        if mode == "train":
            ref.set_jitter(**jit)
            mine.set_jitter(**jit)
        # the reflected passes get their noise per masked ray: run the oracle first to learn its mask, then
        # draw one row per RAY and hand each model the rows of its own masked rays
        out_ref = ref(ref_b) if mode == "eval" else None
        if mode == "train":
            jr = torch.rand(n, sizes["num_reflect_coarse_samples"] + 1, generator=g)
            jp = torch.rand(n, sizes["num_reflect_importance_samples"] + 1, generator=g)
            probe = ref(U.RayBundle(origins=ref_b.origins, directions=ref_b.directions, pixel_area=ref_b.pixel_area))
            mref = probe["mask"]
            ref.set_jitter(**jit, reciprocal=jr[mref], reflect_pdf=jp[mref])
            out_ref = ref(ref_b)
            mine.set_jitter(**jit)
            mmine = mine._get_outputs_nograd(mine.collider(my_b))["mask"].cpu()
            mine.set_jitter(**jit, reciprocal=jr[mmine], reflect_pdf=jp[mmine])
        out = mine._get_outputs_nograd(mine.collider(my_b))
    assert set(out) == set(out_ref)
    mask_ref, mask = out_ref["mask"], out["mask"].cpu()
    agree = mask_ref == mask
    assert agree.float().mean() > 0.97            # borderline rays (acc ~ 1e-2, n.d ~ 0) may flip
    assert mask.any()
    for k in ("mid_rgb_coarse", "mid_rgb_fine", "diff", "tint", "roughness", "accumulation_coarse",
              "accumulation_fine"):
        torch.testing.assert_close(out[k].cpu(), out_ref[k].detach(), rtol=0, atol=ATOL_RGB, msg=lambda s, k=k: f"{k}: {s}")
    for k in ("mid_reflect_coarse", "mid_reflect_fine"):
        torch.testing.assert_close(out[k].cpu()[agree], out_ref[k].detach()[agree], rtol=0, atol=ATOL_RGB,
                                   msg=lambda s, k=k: f"{k}: {s}")
    for k, v in out_ref.items():
        if k not in ("depth_reflect_fine",):
            assert out[k].shape == v.shape, k
    with torch.no_grad():
        loss = mine.get_loss_dict(out, {"image": img.cuda()})
        loss_ref = ref.get_loss_dict({k: v.detach() for k, v in out_ref.items()}, {"image": img})
    assert set(loss) == set(loss_ref)
    for k in ("loss_mid_coarse", "loss_mid_fine"):
        torch.testing.assert_close(loss[k].cpu(), loss_ref[k], rtol=5e-2, atol=1e-4)


def test_state_dict_keys_match_reference_field():
    ref, mine = _models(GOLDEN_SIZES)
    assert sorted(mine.field.state_dict()) == sorted(ref.field.state_dict())
    assert sum(p.numel() for p in mine.field.parameters()) == 618513
    assert list(mine.get_param_groups()) == ["fields"]
