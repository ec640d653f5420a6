"""Fused field forward (K3+K4+K5+K7, tcgen05) vs the oracle field on the same rays, bins and random-init
weights.  Two references: (i) the fp32 oracle -- tolerance of the bf16 MLP (north_star: rendered RGB within
1e-2); (ii) the same oracle with every GEMM operand rounded to bf16 exactly where the kernel rounds --
tight tolerance, catches layout / schedule / swizzle mistakes that a loose bound would hide."""
import pytest
import torch
import torch.nn.functional as F

from oracle import refpath as R
from reflect_sampling_nerf_b200 import ops, packing
from helpers import synthetic_rays

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.bfloat16().float()


def bf16_field(field, mean, cov, dirs, zero_ide=False):
    """Oracle field with GEMM operands rounded to bf16 (fp32 accumulate), as the kernel computes."""
    sd = {k: v.detach() for k, v in field.state_dict().items()}
    lin = lambda x, w, b: _bf(x) @ _bf(sd[w]).T + sd[b]  # noqa: E731
    enc = R.ipe(mean, cov)
    h = enc
    for l in range(8):
        if l == 4:
            h = torch.cat([enc, h], -1)
        h = F.relu(lin(h, f"mlp_base.layers.{l}.weight", f"mlp_base.layers.{l}.bias"))
    head = lambda n: lin(h, f"field_output_{n}.net.weight", f"field_output_{n}.net.bias")  # noqa: E731
    raw = head("density")
    out = {"raw": raw, "density": F.softplus(raw + 0.5)}
    out["pred_normals"] = F.normalize(-F.normalize(head("normals"), dim=-1), dim=-1)
    rr = head("roughness")
    out["rs"], out["rp"] = torch.sigmoid(rr), F.softplus(rr)
    out["diff"], out["tint"] = torch.sigmoid(head("diff")), torch.sigmoid(head("tint"))
    bott = head("bottleneck")
    ide = torch.zeros(*h.shape[:-1], 34) if zero_ide else R.ide(dirs, out["rp"])
    mid_h = F.relu(lin(torch.cat([ide, bott], -1), "mlp_mid.layers.0.weight", "mlp_mid.layers.0.bias"))
    out["mid"] = torch.sigmoid(lin(mid_h, "field_output_mid.net.weight", "field_output_mid.net.bias"))
    out["rgb"] = out["diff"] + out["tint"] * out["mid"]
    return out


def _setup(n, s, seed, kind="uniform", area=3.2e-6):
    torch.manual_seed(seed)
    field = R.OracleField().eval()
    o, d, pa, _ = synthetic_rays(n, seed, pixel_area=area)
    g = torch.Generator().manual_seed(seed + 1)
    if kind == "uniform":
        nears, fars = torch.full((n, 1), 2.0), torch.full((n, 1), 6.0)
    else:
        nears, fars = torch.zeros(n, 1), torch.full((n, 1), 256.0)
    _, bins = R.spaced_bins(nears, fars, s, kind, torch.rand(n, s + 1, generator=g))
    return field, o, d, pa, bins


@pytest.mark.parametrize("n,s,kind,area", [(8, 64, "uniform", 3.2e-6), (37, 24, "uniform", 8.1e-7),
                                          (300, 128, "uniform", 3.2e-6), (64, 64, "reciprocal", 0.05),
                                          (3, 5, "uniform", 3.2e-6)])
def test_field_forward_matches_oracle(n, s, kind, area):
    field, o, d, pa, bins = _setup(n, s, 11 + n, kind, area)
    wblob, bias = packing.pack_field(field.state_dict())
    sigma, feat = ops.field_forward(wblob.cuda(), bias.cuda(), o.cuda(), d.cuda(), pa.cuda(), bins.cuda())
    torch.cuda.synchronize()
    sigma, feat = sigma.cpu(), feat.cpu()
    ex = lambda x: x[:, None, :].expand(n, s, x.shape[-1])  # noqa: E731
    with torch.no_grad():
        mean, cov = R.frustum_gaussian(ex(o), ex(d), bins[:, :-1, None], bins[:, 1:, None], ex(pa))
        mean, cov = R.contract(mean, cov)
        ref = field.point_heads(mean, cov, ex(d), primary=True)
        emu = bf16_field(field, mean, cov, ex(d))
    # (ii) bf16-emulated reference: tight
    torch.testing.assert_close(feat[..., ops.F_RAW_DENSITY], emu["raw"][..., 0], rtol=2e-2, atol=4e-3)
    torch.testing.assert_close(feat[..., ops.F_RGB], emu["rgb"], rtol=0, atol=4e-3)
    torch.testing.assert_close(feat[..., ops.F_DIFF], emu["diff"], rtol=0, atol=3e-3)
    torch.testing.assert_close(feat[..., ops.F_TINT], emu["tint"], rtol=0, atol=3e-3)
    torch.testing.assert_close(feat[..., ops.F_ROUGH_SIGMOID], emu["rs"][..., 0], rtol=0, atol=3e-3)
    # (i) fp32 oracle: the north_star tolerance for the bf16 MLP
    torch.testing.assert_close(feat[..., ops.F_RGB], ref["rgb"], rtol=0, atol=1e-2)
    torch.testing.assert_close(sigma, ref["density"][..., 0], rtol=3e-2, atol=1e-2)
    torch.testing.assert_close(feat[..., ops.F_ROUGH_SOFTPLUS], ref["roughness_softplus"][..., 0], rtol=0, atol=1e-2)
    cosang = (feat[..., ops.F_NORMAL] * ref["pred_normals"]).sum(-1)
    assert cosang.min() > 0.995
    torch.testing.assert_close(feat[..., ops.F_NDOTD], ref["n_dot_d"][..., 0], rtol=0, atol=3e-2)


def test_inf_color_matches_oracle():
    torch.manual_seed(5)
    field = R.OracleField().eval()
    m = 200
    w = F.normalize(torch.randn(m, 3), dim=-1)
    sq = torch.rand(m, 1) * 0.3 + 1e-3
    wblob, bias = packing.pack_field(field.state_dict())
    rgb = ops.field_inf_color(wblob.cuda(), bias.cuda(), w.cuda(), sq.cuda()).cpu()
    with torch.no_grad():
        ref = field.inf_color(w, sq)
    torch.testing.assert_close(rgb, ref, rtol=0, atol=1e-2)


def test_tmem_operand_form_is_bit_identical(monkeypatch):
    """Product forward (A operand of the hidden layers from TMEM) against the shared-memory operand form of the TEST BUILD
    (librsn_b200_dbg.so, RSN_FWD_TS=0; inference launches only -- a training launch always takes the TMEM form): same
    arithmetic in the same order => identical outputs.  Also pins product == test build for the default form, stash and
    aux included, and that repeated launches are bit-identical (no race between the roles of the kernel)."""
    from reflect_sampling_nerf_b200 import _lib
    field, o, d, pa, bins = _setup(37, 24, 5, "uniform", 8.1e-7)      # 888 points = 7 tiles
    wblob, bias = [t.cuda() for t in packing.pack_field(field.state_dict())]
    args = (o.cuda(), d.cuda(), pa.cuda(), bins.cuda())
    nbytes = _lib.lib().rsn_field_stash_bytes(37 * 24)
    zeros = lambda: torch.zeros(nbytes, dtype=torch.uint8, device="cuda")  # noqa: E731  (unwritten mask slots compare equal)

    def run():
        s, f = ops.field_forward(wblob, bias, *args)
        return (s, f) + tuple(ops.field_forward_train(wblob, bias, 0, *args, stash=zeros()))
    monkeypatch.delenv("RSN_FWD_TS", raising=False)
    prod = run()
    try:
        _lib.use_dbg(True)
        dflt = run()
        monkeypatch.setenv("RSN_FWD_TS", "0")
        ss = run()
    finally:
        _lib.use_dbg(False)
    torch.cuda.synchronize()
    for other in (dflt, ss):
        for a, b in zip(prod[:5], other[:5]):
            assert torch.equal(a, b)
        assert torch.equal(prod[5][..., :7], other[5][..., :7])      # aux: 7 of 8 floats per point are defined
    for _ in range(20):                                                # repeated launches of either form: bit-identical
        monkeypatch.setenv("RSN_FWD_TS", "0")
        try:
            _lib.use_dbg(True)
            again_ss = run()
        finally:
            _lib.use_dbg(False)
        again = run()
        torch.cuda.synchronize()
        for ref, got in ((ss, again_ss), (prod, again)):
            for a, b in zip(ref[:5], got[:5]):
                assert torch.equal(a, b)


def test_two_fields_on_two_streams_do_not_share_bias_state():
    """The bias reaches the kernel by pointer (no __constant__ table rewritten per launch): two fields with different
    parameters evaluated concurrently on two streams each get their own results."""
    torch.manual_seed(1)
    fa, fb = R.OracleField().eval(), R.OracleField().eval()
    _, o, d, pa, bins = _setup(2000, 64, 3)
    args = [t.cuda() for t in (o, d, pa, bins)]
    pa_ = [t.cuda() for t in packing.pack_field(fa.state_dict())]
    pb_ = [t.cuda() for t in packing.pack_field(fb.state_dict())]
    ref_a = ops.field_forward(*pa_, *args)[1].clone()
    ref_b = ops.field_forward(*pb_, *args)[1].clone()
    torch.cuda.synchronize()
    assert not torch.equal(ref_a, ref_b)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for _ in range(4):
        with torch.cuda.stream(s1):
            a = ops.field_forward(*pa_, *args)[1]
        with torch.cuda.stream(s2):
            b = ops.field_forward(*pb_, *args)[1]
        outs.append((a, b))
    torch.cuda.synchronize()
    for a, b in outs:
        assert torch.equal(a, ref_a) and torch.equal(b, ref_b)


def test_device_side_ray_count_limits_the_pass():
    """n_rays_dev: the launch is sized for the capacity, rows >= the device count are never written."""
    field, o, d, pa, bins = _setup(300, 32, 21)
    wblob, bias = [t.cuda() for t in packing.pack_field(field.state_dict())]
    args = (o.cuda(), d.cuda(), pa.cuda(), bins.cuda())
    full_s, full_f = ops.field_forward(wblob, bias, *args)
    for m in (0, 1, 77, 300, 1000):
        count = torch.tensor([m], dtype=torch.int32, device="cuda")
        n, s = 300, 32
        sig = torch.full((n, s), -7.0, device="cuda")
        feat = torch.full((n, s, 16), -7.0, device="cuda")
        from reflect_sampling_nerf_b200 import _lib
        _lib.call("rsn_field_forward", wblob.data_ptr(), bias.data_ptr(), 0, args[0].data_ptr(), args[1].data_ptr(),
                  args[2].reshape(-1).data_ptr(), args[3].data_ptr(), n, s, sig.data_ptr(), feat.data_ptr(),
                  count.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        k = min(m, n)
        assert torch.equal(sig[:k], full_s[:k]) and torch.equal(feat[:k], full_f[:k])
        assert bool((sig[k:] == -7.0).all()) and bool((feat[k:] == -7.0).all())


def test_pack_kernel_matches_host_packing():
    """csrc/pack.cu (one launch) against the tensor-op packers of packing.py: identical bytes."""
    torch.manual_seed(9)
    sd = {k: v.cuda() for k, v in R.OracleField().state_dict().items()}
    wblob, bias, wblob_t, wd = ops.pack_field(sd)
    rb, rbias = packing.pack_field(sd)
    rbt, rwd = packing.pack_field_t(sd)
    torch.cuda.synchronize()
    assert torch.equal(wblob, rb) and torch.equal(bias, rbias)
    assert torch.equal(wblob_t, rbt) and torch.equal(wd, rwd)


def test_unpack_kernel_matches_host_unpacking():
    """rsn_unpack_grads (one launch) against packing.unpack_grads on a random gradient blob."""
    offs, shapes, total = ops.wgrad_layout()
    blob = torch.randn(total, device="cuda")
    poffs, ptotal = ops.flat_layout()
    assert ptotal == 618513 - 771                      # every parameter except the unused field_output_low
    flat = torch.full((ptotal,), float("nan"), device="cuda")
    ops.unpack_grads_flat(blob, flat)
    ref = packing.unpack_grads(blob, offs, shapes)
    torch.cuda.synchronize()
    assert not torch.isnan(flat).any()
    for name, off in zip(ops.PACK_ORDER, poffs):
        g = ref[name]
        assert torch.equal(flat[off: off + g.numel()].view(g.shape), g), name
