"""Direct parity of the two encodings (hard-part 4: fp32 trig at |arg| up to 8.6e5):
  * the stand-alone encoders (csrc/encode.cu = the fused kernel's device functions, fp32 out) against oracle.refpath.ipe /
    ide, at an fp32-level tolerance, for means out to the far-256 reflected range and the uncontracted infinity colour;
  * the encoded operand blocks the fused TRAINING forward actually stashed (bf16 block images, un-swizzled here) against
    the same oracle at bf16 half-ulp."""
import pytest
import torch
import torch.nn.functional as F

from helpers import synthetic_rays
from oracle import refpath as R
from reflect_sampling_nerf_b200 import _lib, ops, packing
from reflect_sampling_nerf_b200.components import IntegratedSHEncoding, NeRFEncoding

pytestmark = pytest.mark.gpu


def _unswizzle(block_bytes: torch.Tensor) -> torch.Tensor:
    """[128 rows][128 bytes] 128B-swizzled block image -> [128, 64] bf16 (csrc/umma.cuh block_off)."""
    b = block_bytes.reshape(128, 8, 16)
    rows = torch.arange(128)[:, None]
    chunks = torch.arange(8)[None, :]
    src = chunks ^ (rows & 7)
    out = b[rows, src]                      # logical chunk c of row r lives at physical chunk c ^ (r & 7)
    return out.reshape(128, 128).contiguous().view(torch.bfloat16).reshape(128, 64)


@pytest.mark.parametrize("scale", [0.5, 2.0, 6.6])
def test_ipe_encoder_matches_oracle_fp32(scale):
    """|2 pi x f| reaches 2 pi * 2 * 65536 = 8.2e5 for contracted means (|x| < 2) and beyond for the uncontracted
    infinity-colour mean 2 w; sin() is periodic, so an fp32 argument error of 1 ulp(8e5) = 0.06 rad would be visible
    unless the reduction is exact -- but those frequencies are damped to e^-24 = 0 unless the variance is tiny."""
    g = torch.Generator().manual_seed(int(scale * 10))
    p = 20000
    x = (torch.rand(p, 3, generator=g) * 2 - 1) * scale
    var = torch.exp(torch.rand(p, 3, generator=g) * 30 - 32)            # diag(cov) from 1e-14 to 0.13
    cov = torch.diag_embed(var)
    enc = NeRFEncoding()
    got = enc(x.cuda(), covs=cov.cuda()).cpu()
    ref = R.ipe(x, cov)
    assert got.shape == ref.shape == (p, 99)
    assert torch.equal(got[:, 96:], x)                                   # include_input, appended last
    # where the oracle's own fp32 argument is exact to << 1 rad the features agree to ~1e-6; at the highest frequencies
    # an undamped feature inherits the fp32 rounding of s = fl(fl(2 pi x) f), which oracle and kernel round identically
    err = (got - ref).abs()
    assert float(err.max()) < 2e-6, float(err.max())
    # without covariances: plain sin / cos features (NeRFEncoding.forward(x))
    torch.testing.assert_close(enc(x.cuda()).cpu(), R.ipe(x, torch.zeros(p, 3, 3)), rtol=0, atol=2e-6)


def test_ide_encoder_matches_oracle_fp32():
    g = torch.Generator().manual_seed(2)
    p = 20000
    d = F.normalize(torch.randn(p, 3, generator=g), dim=-1)
    rho = torch.exp(torch.rand(p, 1, generator=g) * 8 - 6)                # softplus roughness from 2e-3 to 7
    got = IntegratedSHEncoding()(d.cuda(), rho.cuda()).cpu()
    ref = R.ide(d, rho)
    assert got.shape == ref.shape == (p, 34)
    # the degree-8 band evaluates polynomials like 6435 z^8 - 12012 z^6 + 6930 z^4 - 1260 z^2 + 35 with heavy cancellation:
    # two fp32 evaluation orders (oracle: z**8 ...; kernel: products of z^2, z^4) differ by up to ~1e-5 absolute there
    torch.testing.assert_close(got[:, :17], ref[:, :17], rtol=2e-5, atol=2e-6)
    torch.testing.assert_close(got[:, 17:], ref[:, 17:], rtol=2e-5, atol=3e-5)


@pytest.mark.parametrize("kind,area,n,s", [("uniform", 3.2e-6, 96, 64), ("reciprocal", 0.02, 64, 64)])
def test_stashed_encoding_blocks_are_the_oracle_encodings_at_bf16_half_ulp(kind, area, n, s):
    """What the MMA consumed: the IPE blocks (stash blocks 0-1) and the IDE block (stash block 38) of the training
    forward, for primary samples and for reflected-style bundles reaching the far plane at 256."""
    torch.manual_seed(3)
    field = R.OracleField().eval()
    o, d, pa, _ = synthetic_rays(n, 5, pixel_area=area)
    g = torch.Generator().manual_seed(6)
    nears, fars = (torch.full((n, 1), 2.0), torch.full((n, 1), 6.0)) if kind == "uniform" else (torch.zeros(n, 1), torch.full((n, 1), 256.0))
    _, bins = R.spaced_bins(nears, fars, s, kind, torch.rand(n, s + 1, generator=g))
    wblob, bias = [t.cuda() for t in packing.pack_field(field.state_dict())]
    sigma, feat, stash, aux = ops.field_forward_train(wblob, bias, 0, o.cuda(), d.cuda(), pa.cuda(), bins.cuda())
    torch.cuda.synchronize()
    tile_bytes = _lib.lib().rsn_field_stash_bytes(128)
    stash = stash.cpu().reshape(-1, tile_bytes)
    ex = lambda x: x[:, None, :].expand(n, s, x.shape[-1])  # noqa: E731
    with torch.no_grad():
        # The Gaussians the fused prologue computed (same device arithmetic: csrc/blob.cu mirrors it operation by operation;
        # tests/test_field_api_gpu.py pins them against the oracle's at 1e-6).  Feeding the ORACLE's own Gaussians instead
        # would add their 1-ulp differences in x (ATen's CPU sqrt / reduction order) times 2 pi f -- up to 1e-3 at the
        # undamped frequencies -- to what is meant to be a test of the encoding and its bf16 rounding.
        mean_g, cov_g = ops.contract(*ops.frustum_gaussians(o.cuda(), d.cuda(), pa.cuda(), bins.cuda()))
        mean_o, cov_o = R.contract(*R.frustum_gaussian(ex(o), ex(d), bins[:, :-1, None], bins[:, 1:, None], ex(pa)))
        torch.testing.assert_close(mean_g.cpu(), mean_o, rtol=2e-6, atol=2e-6)
        ipe = R.ipe(mean_g.cpu(), cov_g.cpu()).reshape(-1, 99)
        ide = R.ide(ex(d), feat[..., ops.F_ROUGH_SOFTPLUS, None].cpu()).reshape(-1, 34)
    n_tiles = (n * s) // 128
    enc_got = torch.cat([torch.cat([_unswizzle(stash[t, 0:16384]), _unswizzle(stash[t, 16384:32768])], dim=1)
                         for t in range(n_tiles)]).float()
    ide_got = torch.cat([_unswizzle(stash[t, 38 * 16384:39 * 16384]) for t in range(n_tiles)]).float()
    assert float(enc_got[:, 99:].abs().max()) == 0.0 and float(ide_got[:, 34:].abs().max()) == 0.0     # zero padding
    # bf16 half-ulp: 8 significant bits => |x_bf16 - x| <= 2^-9 * 2^ceil(log2 |x|) <= 2^-8 |x| (+ the fp32-level error of the
    # encoder itself, up to 3e-5 in the cancelling degree-8 IDE band)
    for got, ref, name in ((enc_got[:, :99], ipe, "ipe"), (ide_got[:, :34], ide, "ide")):
        tol = ref.abs() * 2.0 ** -8 * 1.001 + (4e-6 if name == "ipe" else 3e-5)
        bad = (got - ref).abs() > tol
        assert not bool(bad.any()), (name, int(bad.sum()), float((got - ref).abs().max()))
