"""The reference's method-by-method Field API (reflect_sampling_nerf_field.py:90-207) and component call forms
(reflect_sampling_nerf_model.py:109-124,148-226) on the kernels, against the oracle's restatement of the same methods --
written the way the reference's get_outputs calls them."""
import pytest
import torch
import torch.nn.functional as F
from torch import nn

from helpers import oracle_model, synthetic_rays
from oracle import refpath as R
from oracle import upstream as U
from reflect_sampling_nerf_b200 import components as C
from reflect_sampling_nerf_b200 import ops
from reflect_sampling_nerf_b200.model import ReflectSamplingNeRFModel, ReflectSamplingNeRFModelConfig
from reflect_sampling_nerf_b200.rays import RayBundle

pytestmark = pytest.mark.gpu
SIZES = dict(num_coarse_samples=48, num_importance_samples=48, num_reflect_coarse_samples=24, num_reflect_importance_samples=24)


def _pair(seed=3, train=False):
    ref = oracle_model(SIZES, seed=seed)
    mine = ReflectSamplingNeRFModel(ReflectSamplingNeRFModelConfig(**SIZES)).cuda()
    mine.field.load_state_dict(ref.field.state_dict(), strict=False)
    ref.train(train)
    mine.train(train)
    return ref, mine


def _bundle(n, seed):
    o, d, pa, _ = synthetic_rays(n, seed, pixel_area=3.2e-6)
    nears, fars = torch.full((n, 1), 2.0), torch.full((n, 1), 6.0)
    return (U.RayBundle(origins=o, directions=d, pixel_area=pa, nears=nears, fars=fars),
            RayBundle(origins=o.cuda(), directions=d.cuda(), pixel_area=pa.cuda(), nears=nears.cuda(), fars=fars.cuda()))


def test_reference_call_sequence_on_the_method_api():
    """model.py:148-226, line by line, through sampler(ray_bundle), field.get_blob / contract / get_density / heads /
    get_mid, ray_samples.get_weights and the renderer modules."""
    n = 200
    ref, mine = _pair()
    rb, mb = _bundle(n, 7)
    field, rf = mine.field, ref.field
    with torch.no_grad():
        # uniform sampling (model.py:148): eval mode -> deterministic bins, bit-exact
        samples = mine.sampler_uniform(mb)
        rsamples = ref.sampler_uniform(rb)
        assert torch.equal(samples.frustums.starts.cpu(), rsamples.frustums.starts)
        assert torch.equal(samples.spacing_ends.cpu(), rsamples.spacing_ends)
        assert torch.equal(samples.deltas.cpu(), rsamples.deltas)
        # first pass (model.py:151-156)
        mean, cov = field.get_blob(samples)
        g = rsamples.frustums.get_gaussian_blob()
        torch.testing.assert_close(mean.cpu(), g.mean, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(cov.cpu(), g.cov, rtol=1e-4, atol=1e-12)
        mean_c, cov_c = field.contract(mean, cov)
        rmean_c, rcov_c = R.contract(g.mean, g.cov)
        torch.testing.assert_close(mean_c.cpu(), rmean_c, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(cov_c.cpu(), rcov_c, rtol=1e-3, atol=1e-11)
        density, emb = field.get_density(mean_c, cov_c, True)
        heads = rf.point_heads(rmean_c, rcov_c, rsamples.frustums.directions, primary=True)
        assert density.shape == (n, 48, 1) and tuple(emb.shape) == (n, 48, 256)
        torch.testing.assert_close(density.cpu(), heads["density"], rtol=3e-2, atol=1e-2)
        weights = samples.get_weights(density)
        rweights = rsamples.get_weights(heads["density"])
        torch.testing.assert_close(weights.cpu(), rweights, rtol=0, atol=2e-3)
        acc = mine.renderer_accumulation(weights)
        depth = mine.renderer_depth(weights, samples)
        torch.testing.assert_close(acc.cpu(), U.AccumulationRenderer.forward(rweights), rtol=0, atol=3e-3)
        rdepth = U.DepthRenderer()(rweights, rsamples)
        assert depth.shape == (n, 1)
        assert float(torch.isclose(depth.cpu(), rdepth, rtol=5e-2, atol=0).float().mean()) > 0.9      # a bin mid-point each
        # the renderers on IDENTICAL weights: exact same median bin, sums to fp32 rounding
        wsame = rweights.cuda()
        assert torch.equal(mine.renderer_depth(wsame, samples).cpu(), rdepth)
        torch.testing.assert_close(mine.renderer_accumulation(wsame).cpu(), U.AccumulationRenderer.forward(rweights), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(mine.renderer_normals(heads["pred_normals"].cuda(), wsame).cpu(),
                                   U.NormalsRenderer.forward(heads["pred_normals"], rweights), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(mine.renderer_reflect(heads["rgb"].cuda(), wsame, background_color=torch.rand(n, 3, generator=torch.Generator().manual_seed(1)).cuda()).cpu(),
                                   R.blend(heads["rgb"], rweights, torch.rand(n, 3, generator=torch.Generator().manual_seed(1)), training=False), rtol=1e-5, atol=1e-6)
        # heads (model.py:158-175)
        pn = field.get_pred_normals(emb)
        assert float((pn.cpu() * heads["pred_normals"]).sum(-1).min()) > 0.995
        refl, ndd = field.get_reflection(samples.frustums.directions, pn)
        torch.testing.assert_close(ndd.cpu(), heads["n_dot_d"], rtol=0, atol=3e-2)
        diff, tint = field.get_diff(emb), field.get_tint(emb)
        torch.testing.assert_close(diff.cpu(), heads["diff"], rtol=0, atol=5e-3)
        torch.testing.assert_close(tint.cpu(), heads["tint"], rtol=0, atol=5e-3)
        rough_sp = field.get_roughness(emb, nn.Softplus())
        rough_sg = field.get_roughness(emb)
        torch.testing.assert_close(rough_sp.cpu(), heads["roughness_softplus"], rtol=0, atol=1e-2)
        torch.testing.assert_close(rough_sg.cpu(), heads["roughness_sigmoid"], rtol=0, atol=5e-3)
        torch.testing.assert_close(field.get_roughness(emb, nn.Tanh()).cpu(), torch.tanh(torch.log(torch.expm1(heads["roughness_softplus"]))),
                                   rtol=0, atol=2e-2)
        mid = field.get_mid(samples.frustums.directions, rough_sp.detach(), emb, True)
        torch.testing.assert_close(mid.cpu(), heads["mid"], rtol=0, atol=1e-2)
        rgb = mine.renderer_rgb(diff + tint * mid, weights)
        rrgb = R.blend(heads["rgb"], rweights, U.WHITE, training=False)
        torch.testing.assert_close(rgb.cpu(), rrgb, rtol=0, atol=1e-2)
        # pdf sampling + second pass (model.py:182-190)
        samples_f = mine.sampler_pdf(mb, samples, weights)
        rsamples_f = ref.sampler_pdf(rb, rsamples, weights.cpu())          # same weights in: bit-exact bins out
        assert torch.equal(samples_f.frustums.starts.cpu(), rsamples_f.frustums.starts)
        assert torch.equal(samples_f.frustums.ends.cpu(), rsamples_f.frustums.ends)
        # per-ray renders of the bounce set-up (model.py:215-226)
        tint_r = mine.renderer_factor(tint, weights)
        torch.testing.assert_close(tint_r.cpu(), torch.clamp(torch.sum(rweights * heads["tint"], dim=-2), 0, 1), rtol=0, atol=6e-3)
        nrm_r = mine.renderer_normals(pn, weights)
        assert nrm_r.shape == (n, 3)
        rough_r = mine.renderer_roughness(rough_sg, weights)
        torch.testing.assert_close(rough_r.cpu(), torch.sum(rweights * heads["roughness_sigmoid"], dim=-2), rtol=0, atol=5e-3)
        # a different view direction / roughness at get_mid re-evaluates (IDE input), and get_low = zero IDE
        other = F.normalize(torch.randn(n, 48, 3), dim=-1)
        mid2 = field.get_mid(other.cuda(), rough_sp * 0.5, emb, True)
        ref_mid2 = rf.rgb_from(R.ide(other, heads["roughness_softplus"] * 0.5), heads["embedding"])
        torch.testing.assert_close(mid2.cpu(), ref_mid2, rtol=0, atol=1e-2)
        low = field.get_low(emb)
        torch.testing.assert_close(low.cpu(), rf.rgb_from(torch.zeros(n, 48, 34), heads["embedding"]), rtol=0, atol=1e-2)


def test_get_density_on_arbitrary_gaussians_and_inf_color():
    ref, mine = _pair(seed=5)
    p = 3000
    g = torch.Generator().manual_seed(1)
    mean = torch.randn(p, 3, generator=g) * 0.8
    a = torch.randn(p, 3, 3, generator=g) * 0.02
    cov = a @ a.transpose(-1, -2)
    with torch.no_grad():
        density, emb = mine.field.get_density(mean.cuda(), cov.cuda())
        rd, remb = ref.field.density(mean, cov, want_grad=False)
        torch.testing.assert_close(density.cpu(), rd, rtol=3e-2, atol=1e-2)
        d = F.normalize(torch.randn(p, 3, generator=g), dim=-1)
        rough = mine.field.get_roughness(emb, nn.Softplus())
        mid = mine.field.get_mid(d.cuda(), rough, emb)
        rrough = F.softplus(ref.field.field_output_roughness(remb))
        torch.testing.assert_close(mid.cpu(), ref.field.rgb_from(R.ide(d, rrough), remb), rtol=0, atol=1e-2)
        w = F.normalize(torch.randn(p, 3, generator=g), dim=-1)
        sq = torch.rand(p, 1, generator=g) * 0.3 + 1e-3
        torch.testing.assert_close(mine.field.get_inf_color(w.cuda(), sq.cuda()).cpu(), ref.field.inf_color(w, sq), rtol=0, atol=1e-2)
    with pytest.raises(TypeError, match="embedding handle"):
        mine.field.get_diff(torch.zeros(4, 256).cuda())


def test_get_normals_in_training_mode():
    """field.py:125-127,146-147: get_density(mean, cov, True) in training mode arms get_normals()."""
    n = 64
    ref, mine = _pair(seed=9, train=True)
    rb, mb = _bundle(n, 11)
    mine.sampler_uniform.injected_rand = ref_rand = torch.rand(n, 49, generator=torch.Generator().manual_seed(2))
    ref.set_jitter(uniform=ref_rand)
    samples = mine.sampler_uniform(mb)
    mean, cov = mine.field.contract(*mine.field.get_blob(samples))
    with pytest.raises(RuntimeError, match="get_normals"):
        mine.field.get_normals()
    density, emb = mine.field.get_density(mean, cov, True)
    normals = mine.field.get_normals()
    rsamples = ref.sampler_uniform(rb)
    assert torch.equal(samples.frustums.starts.cpu(), rsamples.frustums.starts)        # stratified, same noise: bit-exact
    g = rsamples.frustums.get_gaussian_blob()
    rmean, rcov = R.contract(g.mean, g.cov)
    heads = ref.field.point_heads(rmean.detach(), rcov.detach(), rsamples.frustums.directions, primary=True)
    cosang = (normals.cpu() * heads["normals"].detach()).sum(-1).flatten()
    q = torch.quantile(cosang, torch.tensor([0.05, 0.5]))
    assert float(q[1]) > 0.999 and float(q[0]) > 0.95, q.tolist()
    assert sum(s.in_flight for s in mine.field._stash_pool) == 0


@pytest.mark.parametrize("tan", [0.25, 1.0, 3.0])
def test_reciprocal_sampler_any_tan_bit_exact(tan):
    """components.py:14-36 with the constructor's `tan` (the model uses 0.25)."""
    n, s = 129, 64
    g = torch.Generator().manual_seed(int(tan * 4))
    rand = torch.rand(n, s + 1, generator=g)
    nears, fars = torch.zeros(n, 1), torch.full((n, 1), 256.0)
    o, d, pa, _ = synthetic_rays(n, 3)
    mine = C.ReciprocalSampler(tan=tan, num_samples=s).train()
    mine.injected_rand = rand
    mb = RayBundle(origins=o.cuda(), directions=d.cuda(), pixel_area=pa.cuda(), nears=nears.cuda(), fars=fars.cuda())
    got = mine(mb)
    # the oracle's own stage function with an explicit noise tensor
    lin = torch.linspace(0.0, 1.0, s + 1)[None]
    centers = (lin[:, 1:] + lin[:, :-1]) / 2.0
    upper, lower = torch.cat([centers, lin[:, -1:]], -1), torch.cat([lin[:, :1], centers], -1)
    bins = lower + (upper - lower) * rand
    fn, inv = (lambda x: x / (1 / tan + x)), (lambda x: x / tan / (1 - x))
    s_near, s_far = fn(nears), fn(fars)
    euclid = inv(bins * s_far + (1 - bins) * s_near)
    assert torch.equal(got.spacing_starts[..., 0].cpu(), bins[:, :-1])
    assert torch.equal(got.frustums.ends[..., 0].cpu(), euclid[:, 1:])
    torch.testing.assert_close(got.spacing_to_euclidean_fn(got.spacing_ends[..., 0]).cpu(), euclid[:, 1:], rtol=1e-6, atol=0)
