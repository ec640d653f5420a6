"""K9 kernels (reflection set-up, masked composition of the bounce) against the reference's per-ray torch formulas
(reflect_sampling_nerf_model.py:215-229, 267-271, 311-313) -- values, mask bits and gradients."""
import pytest
import torch
import torch.nn.functional as F

from reflect_sampling_nerf_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("clamp", [False, True])
def test_reflect_setup_matches_reference_formulas(clamp):
    n = 1000
    g = torch.Generator().manual_seed(3)
    comp = (torch.rand(n, 16, generator=g) * 1.4 - 0.2).cuda()
    acc = torch.rand(n, 1, generator=g).cuda()
    acc[:50] = 0.005
    depth = (2 + 4 * torch.rand(n, 1, generator=g)).cuda()
    d = F.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
    o = torch.randn(n, 3, generator=g).cuda()
    diff, tint, nrm, ndd, mask, o2, wr = ops.reflect_setup(comp, acc, depth, o, d, clamp)
    cl = (lambda x: torch.clamp(x, 0, 1)) if clamp else (lambda x: x)
    torch.testing.assert_close(diff, cl(comp[:, 3:6] + (1 - acc)))
    torch.testing.assert_close(tint, cl(comp[:, 6:9]))
    v = comp[:, 9:12]
    nref = v / (torch.linalg.norm(v, dim=-1, keepdim=True) + 1e-10)
    torch.testing.assert_close(nrm, nref, rtol=1e-5, atol=1e-6)
    nd = torch.sum(nref * d, dim=-1, keepdim=True)
    torch.testing.assert_close(ndd, nd, rtol=1e-5, atol=1e-6)
    ref_mask = torch.logical_and(acc > 1e-2, nd < 0).reshape(-1)
    sure = (nd.abs() > 1e-5).reshape(-1)
    assert torch.equal(mask[sure], ref_mask[sure])
    torch.testing.assert_close(o2, o + depth * d)
    torch.testing.assert_close(wr, F.normalize(d - 2 * nd * nref, dim=-1), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("clamp_inner", [False, True])
def test_reflect_compose_forward_and_backward(clamp_inner):
    n, m = 500, 213
    g = torch.Generator().manual_seed(5)
    base = torch.rand(n, 3, generator=g).cuda().requires_grad_(True)
    diff, tint = torch.rand(n, 3, generator=g).cuda() * 0.6, torch.rand(n, 3, generator=g).cuda()
    idx = torch.sort(torch.randperm(n, generator=g)[:m])[0].cuda()
    comp = (torch.rand(m, 16, generator=g) * 1.2).cuda().requires_grad_(True)
    bg = torch.rand(m, 3, generator=g).cuda().requires_grad_(True)
    acc = torch.rand(m, 1, generator=g).cuda()
    gout = torch.randn(n, 3, generator=g).cuda()
    out = ops.reflect_compose(base, diff, tint, idx, comp, bg, acc, clamp_inner)
    (out * gout).sum().backward()
    got = (out.detach().clone(), base.grad.clone(), comp.grad.clone(), bg.grad.clone())
    base.grad = comp.grad = bg.grad = None
    refl = comp[:, :3] + bg * (1 - acc)
    if clamp_inner:
        refl = torch.clamp(refl, 0, 1)
    ref = base.index_put((idx,), torch.clip(diff[idx] + tint[idx] * refl, 0.0, 1.0))
    torch.testing.assert_close(got[0], ref.detach())
    if not clamp_inner:       # training form: gradients
        (ref * gout).sum().backward()
        torch.testing.assert_close(got[1], base.grad)
        torch.testing.assert_close(got[2], comp.grad)
        torch.testing.assert_close(got[3], bg.grad)


def test_reflect_compose_without_bounced_rays():
    base = torch.rand(7, 3).cuda()
    out = ops.reflect_compose(base, base, base, torch.zeros(0, dtype=torch.int64).cuda(), torch.zeros(0, 16).cuda(),
                              torch.zeros(0, 3).cuda(), torch.zeros(0, 1).cuda())
    assert torch.equal(out, base)
