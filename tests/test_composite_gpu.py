"""K8 parity: weights within 1e-5 relative (fp32), accumulation / features / median depth, and the
hand-written backward against oracle autograd."""
import pytest
import torch

from oracle import refpath as R
from reflect_sampling_nerf_b200 import ops

pytestmark = pytest.mark.gpu
RTOL_W = 1e-5          # north_star: compositing weights within 1e-5 relative in fp32
# alpha = 1 - exp(-sigma*delta) cancels: one ulp of exp (libm-dependent, CUDA expf vs SLEEF) is 6e-8
# ABSOLUTE on alpha, i.e. > 1e-5 relative once alpha < 6e-3.  The reference's own fp32 result carries
# that noise, so the relative bound is paired with an absolute floor of 2 ulp(1.0).
ATOL_W = 1.2e-7


def _inputs(n, S, C, seed, dense=False):
    g = torch.Generator().manual_seed(seed)
    nears, fars = torch.full((n, 1), 2.0), torch.full((n, 1), 6.0)
    _, bins = R.spaced_bins(nears, fars, S, "uniform", torch.rand(n, S + 1, generator=g))
    # sigma ~ softplus(N(0,1)+0.5) * U(0,20)
    sigma = torch.nn.functional.softplus(torch.randn(n, S, generator=g) + 0.5) * torch.rand(n, S, generator=g) * (20 if dense else 3)
    sigma[: n // 8] = 0                       # empty rays
    sigma[n // 8: n // 4, S // 2:] *= 50      # hard surfaces
    feat = torch.rand(n, S, C, generator=g) if C else None
    return sigma, bins.contiguous(), feat


def _oracle(sigma, bins, feat):
    out = R.composite(sigma[..., None], bins[:, :-1, None], bins[:, 1:, None])
    w = out["weights"]
    fo = torch.sum(w * feat, dim=-2) if feat is not None else None
    return w[..., 0], out["accumulation"][..., 0], out["depth"][..., 0], fo


@pytest.mark.parametrize("S", [1, 5, 32, 64, 100, 128, 256, 300])
@pytest.mark.parametrize("C", [0, 1, 3, 16])
def test_composite_forward(S, C):
    sigma, bins, feat = _inputs(301, S, C, seed=S + C)
    w_ref, acc_ref, depth_ref, fo_ref = _oracle(sigma, bins, feat)
    w, acc, depth, fo = ops.composite(sigma.cuda(), bins.cuda(), None if feat is None else feat.cuda())
    torch.testing.assert_close(w.cpu(), w_ref, rtol=RTOL_W, atol=ATOL_W)
    torch.testing.assert_close(acc.cpu(), acc_ref, rtol=1e-5, atol=1e-6)
    if C:
        torch.testing.assert_close(fo.cpu(), fo_ref, rtol=1e-5, atol=1e-6)
    # median depth: the index may legitimately flip where the cumulative weight sits within rounding of 0.5
    cw = torch.cumsum(w_ref, -1)
    safe = ((cw - 0.5).abs() > 1e-5).all(-1)
    assert int(safe.sum()) > 0.9 * safe.numel()
    assert torch.equal(depth.cpu()[safe], depth_ref[safe])


def test_composite_empty():
    w, acc, d, fo = ops.composite(torch.zeros(0, 8).cuda(), torch.zeros(0, 9).cuda(), torch.zeros(0, 8, 3).cuda())
    assert w.shape == (0, 8) and fo.shape == (0, 3)


@pytest.mark.parametrize("S", [7, 64, 128, 200])
@pytest.mark.parametrize("C", [0, 3, 16])
def test_composite_backward(S, C):
    sigma, bins, feat = _inputs(129, S, C, seed=100 + S + C)
    g = torch.Generator().manual_seed(5)
    gw = torch.randn(129, S, generator=g)
    gacc = torch.randn(129, generator=g)
    gfo = torch.randn(129, C, generator=g) if C else None

    s_ref = sigma.clone().requires_grad_(True)
    f_ref = feat.clone().requires_grad_(True) if C else None
    w_ref, acc_ref, _, fo_ref = _oracle(s_ref, bins, f_ref)
    loss = (w_ref * gw).sum() + (acc_ref * gacc).sum() + ((fo_ref * gfo).sum() if C else 0)
    loss.backward()

    s = sigma.cuda().requires_grad_(True)
    f = feat.cuda().requires_grad_(True) if C else None
    w, acc, _, fo = ops.composite(s, bins.cuda(), f)
    loss = (w * gw.cuda()).sum() + (acc * gacc.cuda()).sum() + ((fo * gfo.cuda()).sum() if C else 0)
    loss.backward()
    scale = s_ref.grad.abs().max().item()
    torch.testing.assert_close(s.grad.cpu(), s_ref.grad, rtol=1e-4, atol=1e-5 * scale)
    if C:
        torch.testing.assert_close(f.grad.cpu(), f_ref.grad, rtol=1e-5, atol=3e-7)


def test_composite_full_size_properties():
    """BASELINE C5 coarse pass: 65,536 rays x 128 samples.  Properties: 0 <= w, sum w <= 1, acc == sum w,
    linearity of feat_out in feat."""
    g = torch.Generator(device="cuda").manual_seed(1)
    n, S = 65536, 128
    bins = torch.sort(torch.rand(n, S + 1, device="cuda", generator=g) * 4 + 2, dim=-1).values
    sigma = torch.rand(n, S, device="cuda", generator=g) * 10
    f1 = torch.rand(n, S, 3, device="cuda", generator=g)
    f2 = torch.rand(n, S, 3, device="cuda", generator=g)
    w, acc, depth, o1 = ops.composite(sigma, bins, f1)
    _, _, _, o2 = ops.composite(sigma, bins, f2)
    _, _, _, o12 = ops.composite(sigma, bins, f1 + 2 * f2)
    assert float(w.min()) >= 0 and float(acc.max()) <= 1 + 1e-5
    torch.testing.assert_close(acc, w.sum(-1), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(o12, o1 + 2 * o2, rtol=1e-4, atol=1e-5)
    assert float(depth.min()) >= 2 and float(depth.max()) <= 6


@pytest.mark.parametrize("S", [24, 64, 128, 200])
def test_composite16_fused_normal_losses(S):
    """rsn_composite16_*: 16 channels + the per-sample normal / orientation losses of model.py:403-407 fused in,
    against the plain compositing op followed by the reference's torch formulas (values and gradients)."""
    n = 97
    g = torch.Generator().manual_seed(S)
    sigma = (torch.rand(n, S, generator=g) * 3).cuda().requires_grad_(True)
    bins = (2 + 4 * torch.sort(torch.rand(n, S + 1, generator=g), dim=-1)[0]).cuda()
    feat = torch.rand(n, S, 16, generator=g).cuda()
    feat[..., 0:3] *= 1.6                                      # some composited colours leave [0, 1]: the blend's clip bites
    feat[..., 13] = feat[..., 13] * 2 - 1                      # n.d of either sign
    feat.requires_grad_(True)
    normals = torch.nn.functional.normalize(torch.randn(n, S, 3, generator=g), dim=-1).cuda()
    gw, gacc = torch.rand(n, S, generator=g).cuda(), torch.rand(n, generator=g).cuda()
    gfo = torch.rand(n, 16, generator=g).cuda()
    c_pn, c_ol = 0.3, 0.7

    gblend = torch.randn(n, 3, generator=g).cuda()
    w, acc, depth, fo, pnl, ol, blend = ops.composite16(sigma, bins, feat, normals, blend=True)
    ((w * gw).sum() + (acc * gacc).sum() + (fo * gfo).sum() + c_pn * pnl.sum() + c_ol * ol.sum()
     + (blend * gblend).sum()).backward()
    gs, gf = sigma.grad.clone(), feat.grad.clone()
    sigma.grad = feat.grad = None

    w2, acc2, depth2, fo2 = ops.composite(sigma, bins, feat)
    wd = w2.detach()[..., None]
    pn_loss = torch.sum(wd * torch.sum((normals - feat[..., 9:12]) ** 2, dim=-1, keepdim=True))
    o_loss = torch.sum(wd * torch.clamp_min(feat[..., 13:14], 0.0) ** 2)
    blend2 = torch.clip(fo2[:, 0:3] + (1.0 - acc2[:, None]), 0.0, 1.0)       # renderer_rgb (white) + clip, model.py:176-177
    ((w2 * gw).sum() + (acc2 * gacc).sum() + (fo2 * gfo).sum() + c_pn * pn_loss + c_ol * o_loss
     + (blend2 * gblend).sum()).backward()
    torch.testing.assert_close(blend, blend2, rtol=1e-6, atol=1e-6)
    assert float((blend2 == 1.0).float().mean()) > 0.005         # the clip is active for some rays
    torch.testing.assert_close(w, w2, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(fo, fo2, rtol=1e-5, atol=1e-6)
    assert torch.equal(depth, depth2)
    torch.testing.assert_close(pnl.sum(), pn_loss, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ol.sum(), o_loss, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(gs, sigma.grad, rtol=1e-4, atol=1e-5 * float(sigma.grad.abs().max()))
    torch.testing.assert_close(gf, feat.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("m", [0, 1, 40, 97, 500])
def test_composite16_device_side_ray_count_and_detached_density(m):
    """The bounce form: rows >= the device count are left alone, forward and backward; detach_sigma produces no density
    gradient and the same feature gradient."""
    n, S = 97, 64
    g = torch.Generator().manual_seed(m)
    sigma = (torch.rand(n, S, generator=g) * 3).cuda().requires_grad_(True)
    bins = (2 + 4 * torch.sort(torch.rand(n, S + 1, generator=g), dim=-1)[0]).cuda()
    feat = torch.rand(n, S, 16, generator=g).cuda().requires_grad_(True)
    gfo = torch.rand(n, 16, generator=g).cuda()
    count = torch.tensor([m], dtype=torch.int32, device="cuda")
    k = min(m, n)
    w0, acc0, d0, fo0, _, _, _ = ops.composite16(sigma, bins, feat)
    (fo0[:k] * gfo[:k]).sum().backward()
    gf0 = feat.grad.clone()
    feat.grad = sigma.grad = None
    w, acc, d, fo, _, _, _ = ops.composite16(sigma, bins, feat, None, count, detach_sigma=True)
    assert torch.equal(w[:k], w0[:k]) and torch.equal(fo[:k], fo0[:k]) and torch.equal(d[:k], d0[:k])
    (fo[:k] * gfo[:k]).sum().backward()
    assert sigma.grad is None
    assert torch.equal(feat.grad[:k], gf0[:k])
