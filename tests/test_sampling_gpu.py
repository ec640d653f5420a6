"""K1 / K2 parity (bit-exact bins and indices) -- CUDA through the C-ABI vs oracle.refpath stage oracles."""
import pytest
import torch

from oracle import refpath as R
from reflect_sampling_nerf_b200 import ops

pytestmark = pytest.mark.gpu
KINDS = {"uniform": ops.UNIFORM, "reciprocal": ops.RECIPROCAL}


def _near_far(n, kind, g):
    if kind == "uniform":
        nears = 2.0 + 0.1 * torch.rand(n, 1, generator=g)
        fars = 6.0 + torch.rand(n, 1, generator=g)
    else:  # reflected bundles: near 0, far 256 (model.py:287-288)
        nears = torch.zeros(n, 1)
        fars = torch.full((n, 1), 256.0)
    return nears, fars


@pytest.mark.parametrize("kind", ["uniform", "reciprocal"])
@pytest.mark.parametrize("S", [1, 7, 64, 128, 200])
@pytest.mark.parametrize("jitter", ["none", "full", "single"])
def test_sample_spaced_bit_exact(kind, S, jitter):
    g = torch.Generator().manual_seed(S * 31 + len(kind))
    n = 257
    nears, fars = _near_far(n, kind, g)
    t = None if jitter == "none" else torch.rand(n, S + 1 if jitter == "full" else 1, generator=g)
    ref_s, ref_e = R.spaced_bins(nears, fars, S, kind, t)
    got_s, got_e = ops.sample_spaced(nears.cuda(), fars.cuda(), S, KINDS[kind], None if t is None else t.cuda())
    assert torch.equal(got_s.cpu(), ref_s.contiguous())
    assert torch.equal(got_e.cpu(), ref_e)


def test_sample_spaced_empty():
    s, e = ops.sample_spaced(torch.zeros(0, 1).cuda(), torch.zeros(0, 1).cuda(), 16, ops.UNIFORM)
    assert s.shape == (0, 17) and e.shape == (0, 17)


def _weights(n, S, g, style):
    if style == "peaky":       # a surface: a few large weights
        w = torch.zeros(n, S)
        idx = torch.randint(0, S, (n, 3), generator=g)
        w.scatter_(1, idx, torch.rand(n, 3, generator=g))
        return w / w.sum(-1, keepdim=True).clamp_min(1e-6) * torch.rand(n, 1, generator=g)
    if style == "zero":        # empty rays: exercises the eps padding branch
        return torch.zeros(n, S)
    return torch.rand(n, S, generator=g) * 0.05


@pytest.mark.parametrize("kind", ["uniform", "reciprocal"])
@pytest.mark.parametrize("S,S2", [(64, 64), (128, 128), (24, 24), (100, 37), (256, 128)])
@pytest.mark.parametrize("style", ["peaky", "flat", "zero"])
@pytest.mark.parametrize("train", [True, False])
def test_pdf_resample_bit_exact(kind, S, S2, style, train):
    g = torch.Generator().manual_seed(S * 7 + S2)
    n = 193
    nears, fars = _near_far(n, kind, g)
    t = torch.rand(n, S + 1, generator=g)
    spacing, _ = R.spaced_bins(nears, fars, S, kind, t)
    spacing = spacing.contiguous()
    w = _weights(n, S, g, style)
    rand = torch.rand(n, S2 + 1, generator=g) if train else None
    ref_s, ref_e, ref_i = R.pdf_bins(w, spacing, nears, fars, kind, S2, rand)
    got_s, got_e, got_i = ops.pdf_resample(w.cuda(), spacing.cuda(), nears.cuda(), fars.cuda(), S2, KINDS[kind],
                                           None if rand is None else rand.cuda(), train=train, return_inds=True)
    assert torch.equal(got_i.cpu(), ref_i), "searchsorted indices differ"
    assert torch.equal(got_s.cpu(), ref_s), "spacing bins differ"
    assert torch.equal(got_e.cpu(), ref_e), "euclidean bins differ"
    # property: bins are sorted within [0,1]
    assert bool((got_s[:, 1:] >= got_s[:, :-1]).all()) and float(got_s.min()) >= 0 and float(got_s.max()) <= 1


def test_pdf_resample_full_size_properties():
    """BASELINE C5 sizes: 65,536 rays x 128 -> 128; size-independent properties only."""
    g = torch.Generator(device="cuda").manual_seed(0)
    n, S = 65536, 128
    nears = torch.full((n, 1), 2.0, device="cuda")
    fars = torch.full((n, 1), 6.0, device="cuda")
    sp, _ = ops.sample_spaced(nears, fars, S, ops.UNIFORM, torch.rand(n, S + 1, device="cuda", generator=g))
    w = torch.rand(n, S, device="cuda", generator=g) ** 8
    s2, e2 = ops.pdf_resample(w, sp, nears, fars, S, ops.UNIFORM, train=False)
    assert bool((s2[:, 1:] >= s2[:, :-1]).all())
    assert float(e2.min()) >= 2.0 and float(e2.max()) <= 6.0
    # idempotence of the eval path
    s3, e3 = ops.pdf_resample(w, sp, nears, fars, S, ops.UNIFORM, train=False)
    assert torch.equal(s2, s3) and torch.equal(e2, e3)
