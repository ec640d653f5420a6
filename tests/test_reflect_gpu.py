"""K9 kernels (reflection set-up, device-side compaction, reflected bundle, masked composition of the bounce) against the
oracle (oracle.refpath.reflect_setup = reflect_sampling_nerf_model.py:267-272) and the reference's per-ray torch
formulas (model.py:215-229, 240-241, 311-313) -- values, mask bits, index order and gradients."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import refpath as R
from reflect_sampling_nerf_b200 import ops

pytestmark = pytest.mark.gpu


def _fine_pass(n, seed):
    g = torch.Generator().manual_seed(seed)
    comp = (torch.rand(n, 16, generator=g) * 1.4 - 0.2)
    acc = torch.rand(n, 1, generator=g)
    acc[: n // 20] = 0.005
    depth = 2 + 4 * torch.rand(n, 1, generator=g)
    d = F.normalize(torch.randn(n, 3, generator=g), dim=-1)
    o = torch.randn(n, 3, generator=g)
    return comp, acc, depth, o, d, g


@pytest.mark.parametrize("clamp", [False, True])
def test_reflect_setup_matches_reference_formulas(clamp):
    n = 1000
    comp, acc, depth, o, d, _ = _fine_pass(n, 3)
    diff, tint, nrm, ndd, mask, o2, wr = [t.cpu() for t in ops.reflect_setup(comp.cuda(), acc.cuda(), depth.cuda(), o.cuda(),
                                                                              d.cuda(), clamp)]
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
    assert torch.equal(mask.bool()[sure], ref_mask[sure])
    # the oracle's bounce set-up (model.py:267-272) on every ray
    o2_ref, wr_ref, _ = R.reflect_setup(o, d, depth, nref, nd, comp[:, 12:13])
    torch.testing.assert_close(o2, o2_ref)
    torch.testing.assert_close(wr, wr_ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n", [1, 31, 1000, 4097, 65536])
def test_compaction_is_boolean_indexing_order_without_a_sync(n):
    g = torch.Generator().manual_seed(n)
    mask = torch.rand(n, generator=g) < 0.55
    idx, inv, count = ops.reflect_compact(mask.cuda())
    m = int(mask.sum())
    assert int(count) == m and idx.dtype == torch.int64 and inv.dtype == torch.int32
    assert torch.equal(idx[:m].cpu(), torch.nonzero(mask).reshape(-1))          # what x[mask] would gather, in order
    ref_inv = torch.full((n,), -1, dtype=torch.int32)
    ref_inv[mask] = torch.arange(m, dtype=torch.int32)
    assert torch.equal(inv.cpu(), ref_inv)
    for special in (torch.zeros(n, dtype=torch.bool), torch.ones(n, dtype=torch.bool)):
        idx, inv, count = ops.reflect_compact(special.cuda())
        assert int(count) == int(special.sum())
        if special.all():
            assert torch.equal(idx.cpu(), torch.arange(n))


def test_bundle_matches_oracle_and_its_gradient_reaches_the_roughness_only():
    n = 777
    comp, acc, depth, o, d, g = _fine_pass(n, 9)
    comp_c = comp.cuda().requires_grad_(True)
    diff, tint, nrm, ndd, mask, o2_all, wr_all = ops.reflect_setup(comp_c, acc.cuda(), depth.cuda(), o.cuda(), d.cuda(), False)
    idx, inv, count = ops.reflect_compact(mask)
    m = int(count)
    o2, wr, sqr, area = ops.reflect_bundle(comp_c, idx, inv, count, o2_all, wr_all, ndd)
    mk = mask.bool().cpu()
    rough = comp[:, 12:13].clone().requires_grad_(True)
    o2_ref, wr_ref, sqr_ref = R.reflect_setup(o[mk], d[mk], depth[mk], nrm.cpu()[mk], ndd.cpu()[mk], rough[mk])
    assert torch.equal(o2[:m].cpu(), o2_all.cpu()[mk]) and torch.equal(wr[:m].cpu(), wr_all.cpu()[mk])
    torch.testing.assert_close(o2[:m].cpu(), o2_ref)
    assert torch.equal(sqr[:m].cpu(), sqr_ref.detach()[:, 0])                 # same fp32 operation order: bit-exact
    assert torch.equal(area[:m].cpu(), (math.pi * sqr_ref.detach())[:, 0])    # pixel_area = pi sqradius (model.py:286)
    gs, ga = torch.randn(n, generator=g), torch.randn(n, generator=g)
    ((sqr * gs.cuda())[:m].sum() + (area * ga.cuda())[:m].sum()).backward()
    ((sqr_ref[:, 0] * gs[:m]).sum() + (math.pi * sqr_ref[:, 0] * ga[:m]).sum()).backward()
    got = comp_c.grad.cpu()
    torch.testing.assert_close(got[:, 12], rough.grad[:, 0], rtol=1e-5, atol=1e-7)
    got[:, 12] = 0
    assert float(got.abs().max()) == 0.0


@pytest.mark.parametrize("clamp_inner", [False, True])
def test_reflect_compose_forward_and_backward(clamp_inner):
    n = 500
    g = torch.Generator().manual_seed(5)
    mask = torch.rand(n, generator=g) < 0.4
    idx, inv, count = ops.reflect_compact(mask.cuda())
    m = int(count)
    acc_f = torch.rand(n, generator=g).cuda().requires_grad_(True)
    diff, tint = torch.rand(n, 3, generator=g).cuda() * 0.6, torch.rand(n, 3, generator=g).cuda()
    comp = (torch.rand(n, 16, generator=g) * 1.2).cuda().requires_grad_(True)       # capacity rows; the first m are live
    bg = torch.rand(n, 3, generator=g).cuda().requires_grad_(True)
    acc_r = torch.rand(n, generator=g).cuda()
    depth_r = torch.rand(n, generator=g).cuda() * 50
    gout = torch.randn(n, 3, generator=g).cuda()
    out, depth_pad = ops.reflect_compose(acc_f, diff, tint, inv, comp, bg, acc_r, depth_r, clamp_inner)
    (out * gout).sum().backward()
    got = (out.detach().clone(), acc_f.grad.clone(), comp.grad.clone(), bg.grad.clone())
    acc_f.grad = comp.grad = bg.grad = None
    # reference formulas: white (1 - acc_fine) fallback (model.py:240-241), masked overwrite + clip (model.py:311-313)
    base = torch.ones(n, 3, device="cuda") * (1.0 - acc_f[:, None])
    refl = comp[:m, :3] + bg[:m] * (1 - acc_r[:m, None])
    if clamp_inner:
        refl = torch.clamp(refl, 0, 1)
    sel = idx[:m]
    ref = base.index_put((sel,), torch.clip(diff[sel] + tint[sel] * refl, 0.0, 1.0))
    torch.testing.assert_close(got[0], ref.detach())
    dref = torch.zeros(n, device="cuda").index_put((sel,), depth_r[:m])
    assert torch.equal(depth_pad, dref)                                          # padded depth_reflect_fine
    if not clamp_inner:       # training form: gradients
        (ref * gout).sum().backward()
        torch.testing.assert_close(got[1], acc_f.grad)
        torch.testing.assert_close(got[2][:m], comp.grad[:m])
        torch.testing.assert_close(got[3][:m], bg.grad[:m])


def test_reflect_compose_without_bounced_rays():
    n = 7
    acc = torch.rand(n).cuda()
    idx, inv, count = ops.reflect_compact(torch.zeros(n, dtype=torch.bool).cuda())
    z = torch.zeros(n, 3).cuda()
    out, depth = ops.reflect_compose(acc, z, z, inv, torch.zeros(n, 16).cuda(), z, torch.zeros(n).cuda(), torch.ones(n).cuda())
    assert torch.equal(out, (1 - acc)[:, None].expand(n, 3)) and float(depth.abs().max()) == 0.0
