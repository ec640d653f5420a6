"""Micro-benchmark of the K8 compositing kernels against the HBM roofline (CUDA events; buffers rotate so the working
set exceeds L2).  usage: python scripts/bench_composite.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflect_sampling_nerf_b200 import ops

PEAK = 6543.7


def run(n, s, c, iters=20, nbuf=4):
    torch.manual_seed(0)
    sig = [torch.rand(n, s, device="cuda") * 5 for _ in range(nbuf)]
    bins = [(2 + 4 * torch.linspace(0, 1, s + 1, device="cuda"))[None].expand(n, s + 1).contiguous() for _ in range(nbuf)]
    feat = [torch.rand(n, s, c, device="cuda") for _ in range(nbuf)] if c else [None] * nbuf
    for i in range(3):
        ops.composite(sig[i % nbuf], bins[i % nbuf], feat[i % nbuf])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ops.composite(sig[i % nbuf], bins[i % nbuf], feat[i % nbuf])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fwd_bytes = n * s * (4 + 4 + 4 * c + 4) + n * (s + 1) * 0 + n * 4 * (c + 2) + n * 4      # sigma, bins (~4/sample), feat in; w out; per ray
    print(f"composite_fwd N={n} S={s} C={c}: {ms*1e3:8.1f} us  {fwd_bytes/ms/1e6:7.1f} GB/s ({fwd_bytes/ms/1e6/PEAK*100:5.1f}% of {PEAK})")
    # backward
    outs = []
    for i in range(nbuf):
        sg = sig[i].clone().requires_grad_(True)
        ft = feat[i].clone().requires_grad_(True) if c else None
        w, acc, depth, fo = ops.composite(sg, bins[i], ft)
        outs.append((sg, ft, w, acc, fo))
    gw = torch.rand(n, s, device="cuda"); ga = torch.rand(n, device="cuda"); gf = torch.rand(n, max(c, 1), device="cuda")[:, :c]
    def bwd(i):
        sg, ft, w, acc, fo = outs[i % nbuf]
        loss_out = [w, acc] + ([fo] if c else [])
        grads = [gw, ga] + ([gf] if c else [])
        torch.autograd.backward(loss_out, grads, retain_graph=True)
        sg.grad = None
        if ft is not None: ft.grad = None
    for i in range(3): bwd(i)
    torch.cuda.synchronize()
    e0.record()
    for i in range(iters): bwd(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    bwd_bytes = n * s * (4 + 4 + 4 * c + 4 + 4 + 4 * c) + n * 4 * (c + 2)
    print(f"composite_bwd N={n} S={s} C={c}: {ms*1e3:8.1f} us  {bwd_bytes/ms/1e6:7.1f} GB/s ({bwd_bytes/ms/1e6/PEAK*100:5.1f}% of {PEAK})  (includes autograd glue)")


for n, s in ((16384, 128), (65536, 128), (65536, 256)):
    for c in (3, 16):
        run(n, s, c)
