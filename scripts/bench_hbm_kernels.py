"""Kernel-only micro-benchmark of the HBM-bound stages (K1 spaced sampling, K2 PDF resampling, K8 compositing forward /
backward in the model's 16-channel form) against the measured HBM copy peak: direct C-ABI calls on pre-allocated buffers,
CUDA events, buffer sets rotated so that the working set exceeds L2.
usage: python scripts/bench_hbm_kernels.py [n_rays ...]      (default: 16384 = C2 per GPU, 65536 = C5)"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from reflect_sampling_nerf_b200 import _lib, ops  # noqa: E402

pk = os.path.join(REPO, "MEASURED_PEAKS.json")
PEAK = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6543.7
S, ITERS = 128, 20


def timeit(fn, nbuf):
    for i in range(nbuf):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(ITERS):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / ITERS


def report(name, n, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print(f"{name:34s} N={n:6d} S={S}: {ms * 1e3:8.1f} us  {gbs:7.1f} GB/s  {gbs / PEAK * 100:5.1f} % of {PEAK:.0f}")


def run(n):
    dev = "cuda"
    nbuf = max(2, int(300e6 // (n * S * 160)) + 1)
    g = torch.Generator(device=dev).manual_seed(0)
    nears, fars = torch.full((n,), 2.0, device=dev), torch.full((n,), 6.0, device=dev)
    rnd = [torch.rand(n, S + 1, device=dev, generator=g) for _ in range(nbuf)]
    lin = torch.linspace(0, 1, S + 1, device=dev)
    sp_o, eu_o = torch.empty(n, S + 1, device=dev), torch.empty(n, S + 1, device=dev)
    ms = timeit(lambda i: _lib.call("rsn_sample_spaced", nears.data_ptr(), fars.data_ptr(), lin.data_ptr(), rnd[i].data_ptr(),
                                    S + 1, 0, 0.25, sp_o.data_ptr(), eu_o.data_ptr(), n, S, None, _lib.stream()), nbuf)
    report("K1 sample_spaced", n, ms, n * (S + 1) * 12)
    sp, eu = ops.sample_spaced(nears, fars, S, 0, rnd[0])
    sig = [torch.rand(n, S, device=dev, generator=g) * 5 for _ in range(nbuf)]
    feat = [torch.rand(n, S, 16, device=dev, generator=g) for _ in range(nbuf)]
    nrm = [torch.nn.functional.normalize(torch.randn(n, S, 3, device=dev, generator=g), dim=-1) for _ in range(nbuf)]
    w0 = ops.composite16(sig[0], eu, feat[0])[0]
    wts = [w0 * (0.5 + 0.1 * i) for i in range(nbuf)]
    u_base = ops._pdf_u_base(S, True, torch.device(dev))
    ms = timeit(lambda i: _lib.call("rsn_pdf_resample", wts[i].data_ptr(), S, sp.data_ptr(), nears.data_ptr(), fars.data_ptr(),
                                    u_base.data_ptr(), rnd[i].data_ptr(), 0, 0.25, 0.01, sp_o.data_ptr(), eu_o.data_ptr(), None,
                                    n, S, S, None, _lib.stream()), nbuf)
    report("K2 pdf_resample", n, ms, n * S * 20)
    # K8 forward / backward, direct calls (the model's form: 16 channels + per-sample normal losses + white blend)
    weights = torch.empty(n, S, device=dev)
    acc, depth, pnl, ol = (torch.empty(n, device=dev) for _ in range(4))
    fo, rgb = torch.empty(n, 16, device=dev), torch.empty(n, 3, device=dev)

    def fwd(i):
        _lib.call("rsn_composite16_fwd", sig[i].data_ptr(), eu.data_ptr(), eu.data_ptr() + 4, S + 1, feat[i].data_ptr(),
                  nrm[i].data_ptr(), weights.data_ptr(), acc.data_ptr(), depth.data_ptr(), fo.data_ptr(), pnl.data_ptr(),
                  ol.data_ptr(), rgb.data_ptr(), n, S, None, _lib.stream())
    ms = timeit(fwd, nbuf)
    report("K8 composite16_fwd (+normals)", n, ms, n * S * (4 + 4 + 64 + 12 + 4) + n * 4 * 24)
    g_w, g_fo = torch.rand(n, S, device=dev, generator=g), torch.rand(n, 16, device=dev, generator=g)
    g_acc, g_pnl, g_ol = (torch.rand(n, device=dev, generator=g) for _ in range(3))
    g_rgb = torch.rand(n, 3, device=dev, generator=g)
    g_sig, g_feat = [torch.empty(n, S, device=dev) for _ in range(nbuf)], [torch.empty(n, S, 16, device=dev) for _ in range(nbuf)]

    def bwd(i):
        _lib.call("rsn_composite16_bwd", sig[i].data_ptr(), eu.data_ptr(), eu.data_ptr() + 4, S + 1, feat[i].data_ptr(),
                  nrm[i].data_ptr(), g_w.data_ptr(), g_acc.data_ptr(), g_fo.data_ptr(), g_pnl.data_ptr(), g_ol.data_ptr(),
                  g_rgb.data_ptr(), fo.data_ptr(), acc.data_ptr(), g_sig[i].data_ptr(), g_feat[i].data_ptr(), n, S, None,
                  _lib.stream())
    ms = timeit(bwd, nbuf)
    report("K8 composite16_bwd (+normals)", n, ms, n * S * (4 + 4 + 64 + 12 + 4 + 4 + 64) + n * 4 * 24)


for n in ([int(a) for a in sys.argv[1:]] or [16384, 65536]):
    run(n)
