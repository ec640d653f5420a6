"""Whole training step: get_outputs + get_loss_dict + backward of the drop-in model (hand-written forward,
normals, dgrad and wgrad kernels) against the oracle model's autograd on the same rays, weights and injected
stratification noise.  Gradients are compared parameter by parameter (cosine + norm): the MLP runs in bf16."""
import pytest
import torch
import torch.nn.functional as F

from helpers import oracle_model, synthetic_rays
from oracle import upstream as U
from reflect_sampling_nerf_b200.model import ReflectSamplingNeRFModel, ReflectSamplingNeRFModelConfig
from reflect_sampling_nerf_b200.rays import RayBundle

pytestmark = pytest.mark.gpu
SIZES = dict(num_coarse_samples=32, num_importance_samples=32, num_reflect_coarse_samples=16,
             num_reflect_importance_samples=16)


def _cos(a, b):
    return float(F.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0))


@pytest.mark.parametrize("boost_normal_losses", [False, True])
def test_train_step_gradients_match_oracle(boost_normal_losses):
    n = 512
    ref = oracle_model(SIZES, seed=11).train()
    mine = ReflectSamplingNeRFModel(ReflectSamplingNeRFModelConfig(**SIZES)).cuda().train()
    mine.field.load_state_dict(ref.field.state_dict(), strict=False)
    if boost_normal_losses:      # make the normal / orientation terms dominate so their paths are really tested
        for m in (ref.loss_coefficients, mine.config.loss_coefficients):
            for k in ("predicted_normal_loss_coarse", "predicted_normal_loss_fine", "orientation_loss_coarse",
                      "orientation_loss_fine"):
                m[k] = m[k] * 300.0
    o, d, pa, img = synthetic_rays(n, 23, pixel_area=3.2e-6)
    g = torch.Generator().manual_seed(5)
    jit = dict(uniform=torch.rand(n, SIZES["num_coarse_samples"] + 1, generator=g),
               pdf=torch.rand(n, SIZES["num_importance_samples"] + 1, generator=g))
    jr = torch.rand(n, SIZES["num_reflect_coarse_samples"] + 1, generator=g)
    jp = torch.rand(n, SIZES["num_reflect_importance_samples"] + 1, generator=g)
    # reflected-pass noise is one row per RAY, handed to each model for its own masked rays
    ref.set_jitter(**jit)
    mref = ref(U.RayBundle(origins=o, directions=d, pixel_area=pa))["mask"]
    ref.set_jitter(**jit, reciprocal=jr[mref], reflect_pdf=jp[mref])
    out_ref = ref(U.RayBundle(origins=o, directions=d, pixel_area=pa))
    loss_ref = ref.get_loss_dict(out_ref, {"image": img})
    ref.zero_grad()
    sum(loss_ref.values()).backward()

    bundle = lambda: RayBundle(origins=o.cuda(), directions=d.cuda(), pixel_area=pa.cuda())  # noqa: E731
    mine.set_jitter(**jit)
    with torch.no_grad():
        mmine = mine(bundle())["mask"].cpu()
    assert (mmine == mref).float().mean() > 0.97
    mine.set_jitter(**jit, reciprocal=jr[mmine], reflect_pdf=jp[mmine])
    out = mine(bundle())
    assert set(out) == set(out_ref)
    for k, v in out_ref.items():
        if k != "depth_reflect_fine":
            assert out[k].shape == v.shape, k
            assert out[k].requires_grad == v.requires_grad, k        # same detach topology (App. D)
    loss = mine.get_loss_dict(out, {"image": img.cuda()})
    sum(loss.values()).backward()
    torch.cuda.synchronize()
    for k in loss_ref:
        torch.testing.assert_close(loss[k].detach().cpu(), loss_ref[k].detach(), rtol=6e-2, atol=1e-5,
                                   msg=lambda s, k=k: f"{k}: {s}")
    # density-gradient normals of the training path
    cosang = (out["normals_fine"].cpu() * out_ref["normals_fine"]).sum(-1)
    assert float(cosang.median()) > 0.995
    worst = []
    for (name, p), (_, q) in zip(ref.field.named_parameters(), mine.field.named_parameters()):
        if "field_output_low" in name:
            assert q.grad is None and p.grad is None
            continue
        rg, mg = p.grad, q.grad.cpu()
        c, ratio = _cos(mg, rg), float(mg.norm() / rg.norm())
        worst.append((c, ratio, name))
    print("\n".join(f"{c:.4f} {r:.3f} {nme}" for c, r, nme in sorted(worst)))
    for c, ratio, name in worst:
        assert c > 0.97, (name, c, ratio)
        assert abs(ratio - 1) < 0.1, (name, c, ratio)
