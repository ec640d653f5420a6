"""Micro-benchmark of the fused field forward kernel (CUDA events, inputs resident in HBM).
usage: python scripts/bench_field.py [n_rays] [n_samples] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflect_sampling_nerf_b200 import ops, packing  # noqa: E402
from reflect_sampling_nerf_b200.plugin_field_init import random_field_state  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
s = int(sys.argv[2]) if len(sys.argv) > 2 else 128
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
torch.manual_seed(0)
sd = random_field_state()
wblob, bias = packing.pack_field(sd)
wblob, bias = wblob.cuda(), bias.cuda()
d = torch.nn.functional.normalize(torch.randn(n, 3), dim=-1)
o = -4 * d + 0.3 * torch.randn(n, 3)
pa = torch.full((n,), 3.2e-6)
bins = (2.0 + 4.0 * torch.linspace(0, 1, s + 1))[None].expand(n, s + 1).contiguous()
o, d, pa, bins = o.cuda(), d.cuda(), pa.cuda(), bins.cuda()
for _ in range(3):
    ops.field_forward(wblob, bias, o, d, pa, bins)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.field_forward(wblob, bias, o, d, pa, bins)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
flop = n * s * 1230592
print(f"field_forward N={n} S={s}: {ms:.3f} ms  {n*s/ms/1e3:.1f} Mpts/s  {flop/ms/1e9:.1f} TFLOP/s "
      f"({flop/ms/1e9/1611.1*100:.1f}% of measured bf16 burst peak 1611.1)")
