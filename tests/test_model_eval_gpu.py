"""Whole-path parity: the drop-in model's get_outputs (CUDA kernels through the C-ABI) against the oracle
model on the same rays, random-init weights and (in train mode) injected stratification noise -- at small sizes, at
BASELINE C1's sample counts (64+64 / 64+64, 1,024 rays) and at C2's (128+128 / 64+64).
north_star tolerance: rendered RGB within 1e-2 max-abs for the bf16 MLP."""
import pytest
import torch

from helpers import GOLDEN_SIZES, oracle_model, synthetic_rays
from oracle import upstream as U
from reflect_sampling_nerf_b200.model import ReflectSamplingNeRFModel, ReflectSamplingNeRFModelConfig
from reflect_sampling_nerf_b200.rays import RayBundle

pytestmark = pytest.mark.gpu
ATOL_RGB = 1e-2
C1 = dict(num_coarse_samples=64, num_importance_samples=64, num_reflect_coarse_samples=64, num_reflect_importance_samples=64)
C2 = dict(num_coarse_samples=128, num_importance_samples=128, num_reflect_coarse_samples=64, num_reflect_importance_samples=64)
MID = dict(num_coarse_samples=64, num_importance_samples=64, num_reflect_coarse_samples=32, num_reflect_importance_samples=32)


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
@pytest.mark.parametrize("sizes,n", [(GOLDEN_SIZES, 384), (MID, 384), (C1, 1024), (C2, 256)],
                         ids=["golden", "mid", "C1", "C2"])
def test_get_outputs_matches_oracle(mode, sizes, n):
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
            with torch.no_grad():
                mmine = mine(my_b)["mask"].cpu()
            mine.set_jitter(**jit, reciprocal=jr[mmine], reflect_pdf=jp[mmine])
        with torch.no_grad():
            out = mine(my_b)
    assert set(out) == set(out_ref)
    mask_ref, mask = out_ref["mask"], out["mask"].cpu()
    agree = mask_ref == mask
    assert agree.float().mean() > 0.985            # borderline rays (acc ~ 1e-2, n.d ~ 0) may flip
    assert mask.any()
    assert int(mine.last_num_bounced) == int(mask.sum())
    for k in ("mid_rgb_coarse", "mid_rgb_fine", "diff", "tint", "roughness", "accumulation_coarse",
              "accumulation_fine"):
        torch.testing.assert_close(out[k].cpu(), out_ref[k].detach(), rtol=0, atol=ATOL_RGB, msg=lambda s, k=k: f"{k}: {s}")
    # the bounce: every ray on which the two masks agree (bounced or not) is compared -- nothing is dropped but the flips
    for k in ("mid_reflect_coarse", "mid_reflect_fine"):
        torch.testing.assert_close(out[k].cpu()[agree], out_ref[k].detach()[agree], rtol=0, atol=ATOL_RGB,
                                   msg=lambda s, k=k: f"{k}: {s}")
    for k, v in out_ref.items():
        if k not in ("depth_reflect_fine",):
            assert out[k].shape == v.shape, k
    # depth_reflect_fine: [N,1], zero where no bounce; on the bounced rays = the reference's ragged [M,1] tensor.  The
    # median depth is a bin mid-point: a neighbouring bin is the only way to differ.
    dpad = out["depth_reflect_fine"].cpu()
    assert dpad.shape == (n, 1) and float(dpad[~mask].abs().max() if (~mask).any() else 0.0) == 0.0
    both = mask & mask_ref
    mine_d = dpad[both][:, 0]
    ref_d = torch.zeros(n, 1).masked_scatter(mask_ref[:, None], out_ref["depth_reflect_fine"].detach())[both][:, 0]
    close = torch.isclose(mine_d, ref_d, rtol=2e-2, atol=1e-3)
    assert close.float().mean() > 0.9, float(close.float().mean())
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
    # a reference checkpoint carries the torchmetrics LPIPS network: those keys are ignored, not fatal
    sd = dict(mine.state_dict())
    sd["lpips.net.slice1.0.weight"] = torch.zeros(3)
    mine.load_state_dict(sd, strict=True)


def test_eval_forward_does_not_synchronise_with_the_host():
    """The path never reads the number of bouncing rays (or anything else) back: with sync debugging set to "error" any
    implicit device->host synchronisation inside get_outputs raises."""
    _, mine = _models(GOLDEN_SIZES)
    mine.eval()
    _, my_b, _ = _bundles(256, 5, 3.2e-6)
    with torch.no_grad():
        mine(my_b)                                  # warm-up: lazily built tables, packed weights
        torch.cuda.synchronize()
        torch.cuda.set_sync_debug_mode("error")
        try:
            out = mine(RayBundle(origins=my_b.origins, directions=my_b.directions, pixel_area=my_b.pixel_area))
        finally:
            torch.cuda.set_sync_debug_mode("default")
    torch.cuda.synchronize()
    assert out["mid_reflect_fine"].shape == (256, 3)
