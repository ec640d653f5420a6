"""The fused optimizer (csrc/optim.cu) against torch.optim.RAdam + the reference's exponential decay
(reflect_sampling_nerf_config.py:50-53), and the fused loss kernel (csrc/loss.cu, §8 a19) against the reference's torch
formulas (reflect_sampling_nerf_model.py:395-429)."""
import math

import pytest
import torch

from reflect_sampling_nerf_b200 import ops
from reflect_sampling_nerf_b200.field import ReflectSamplingNeRFNerfField
from reflect_sampling_nerf_b200.optim import FusedRAdam

pytestmark = pytest.mark.gpu


def _fields():
    torch.manual_seed(0)
    a = ReflectSamplingNeRFNerfField().cuda()
    torch.manual_seed(0)
    b = ReflectSamplingNeRFNerfField().cuda()
    return a, b


@pytest.mark.parametrize("decay", [False, True])
def test_fused_radam_matches_torch_radam_over_100_steps(decay):
    fa, fb = _fields()
    offs, total = ops.flat_layout()
    lr, lr_final, max_steps = 1e-3, 1e-4, 60
    opt_t = torch.optim.RAdam([p for p in fb.parameters()], lr=lr, eps=1e-15)
    sched = None
    if decay:      # nerfstudio ExponentialDecayScheduler without warm-up: lr_init^(1-t) lr_final^t, t = clip(step / max_steps)
        f = lambda step: math.exp(math.log(lr) * (1 - min(step / max_steps, 1)) + math.log(lr_final) * min(step / max_steps, 1)) / lr  # noqa: E731
        sched = torch.optim.lr_scheduler.LambdaLR(opt_t, lr_lambda=f)
    opt_f = FusedRAdam(list(fa.parameters()), lr=lr, eps=1e-15, lr_final=lr_final if decay else 0.0,
                       max_steps=max_steps if decay else 0, field=fa)
    named_a, named_b = dict(fa.named_parameters()), dict(fb.named_parameters())
    flat = torch.empty(total, device="cuda")
    fa.__dict__["_flat_grad"] = flat
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(100):
        # heavy-tailed synthetic gradients with a drifting scale (exercises both branches of the rectification: rho_t <= 5
        # for the first 5 steps)
        flat.copy_(torch.randn(total, device="cuda", generator=g) ** 3 * (1e-3 * (1 + step % 7)))
        for k, off in zip(ops.PACK_ORDER, offs):
            named_b[k].grad = flat[off: off + named_b[k].numel()].view_as(named_b[k]).clone()
        opt_f.step()
        opt_t.step()
        if sched is not None:
            sched.step()
    torch.cuda.synchronize()
    assert int(opt_f.steps_taken) == 100
    worst = 0.0
    for k in ops.PACK_ORDER:
        pa, pb = named_a[k].detach(), named_b[k].detach()
        worst = max(worst, float(((pa - pb).abs() / (pb.abs() + 1e-3)).max()))
    assert worst < 1e-6 * 50, worst                     # 100 steps of 1e-3: total drift below 5e-5 relative-to-(|p|+1e-3)
    torch.testing.assert_close(named_a["mlp_base.layers.4.weight"], named_b["mlp_base.layers.4.weight"], rtol=1e-6, atol=1e-6)
    assert torch.equal(fa.field_output_low.net.weight, fb.field_output_low.net.weight)      # never trained (App. B Q18)
    if decay:
        want = lr * math.exp(math.log(lr_final / lr) * 1.0)           # step 99 >= max_steps: clipped at lr_final
        assert abs(float(opt_f._state_dev["lr"]) - want) < 1e-9
    # the re-pack that follows every fused step: operand blobs hold the NEW weights
    wblob, bias = fa.packed()
    ref_blob, ref_bias, _, _ = ops.pack_field(named_a)
    assert torch.equal(wblob, ref_blob) and torch.equal(bias, ref_bias)


def test_fused_loss_matches_reference_formulas_forward_and_backward():
    n = 5000
    g = torch.Generator().manual_seed(3)
    preds = [torch.rand(n, 3, generator=g).cuda().requires_grad_(True) for _ in range(4)]
    image = torch.rand(n, 3, generator=g).cuda()
    sums = [(torch.rand(n, generator=g) * 0.1).cuda().requires_grad_(True) for _ in range(4)]     # pnl_c, pnl_f, ol_c, ol_f
    coef = torch.tensor([1.0, 1.0, 1.0, 1.0, 3e-5, 3e-4, 1e-2, 1e-1], device="cuda")
    ws = ops.loss_workspace("cuda")
    for round_ in range(2):          # twice: the workspace counter must be left clean
        terms, total = ops.fused_loss(*preds, image, *sums, coef, ws)
        mse = torch.nn.MSELoss()
        ref_terms = torch.stack([mse(image, p) for p in preds] + [s.sum() for s in sums]) * coef
        ref_total = ref_terms.sum()
        torch.testing.assert_close(terms, ref_terms.detach(), rtol=1e-5, atol=1e-9)
        torch.testing.assert_close(total, ref_total.detach(), rtol=1e-5, atol=1e-9)
        gw = torch.rand(8, generator=g).cuda()
        ((terms * gw).sum() + 2.0 * total).backward()
        got = [t.grad.clone() for t in preds + sums]
        for t in preds + sums:
            t.grad = None
        ((ref_terms * gw).sum() + 2.0 * ref_total).backward()
        for a, t in zip(got, preds + sums):
            torch.testing.assert_close(a, t.grad, rtol=1e-5, atol=1e-12)
            t.grad = None
