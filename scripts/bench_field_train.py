"""Micro-benchmark of the four field kernels of one training pass (CUDA events, inputs resident in HBM).
usage: python scripts/bench_field_train.py [n_rays] [n_samples] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflect_sampling_nerf_b200 import _lib, ops, packing  # noqa: E402
from reflect_sampling_nerf_b200.plugin_field_init import random_field_state  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
s = int(sys.argv[2]) if len(sys.argv) > 2 else 128
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
torch.manual_seed(0)
sd = random_field_state()
wblob, bias = [t.cuda() for t in packing.pack_field(sd)]
wblob_t, wd = [t.cuda() for t in packing.pack_field_t(sd)]
d = torch.nn.functional.normalize(torch.randn(n, 3), dim=-1)
o = -4 * d + 0.3 * torch.randn(n, 3)
pa = torch.full((n,), 3.2e-6)
bins = (2.0 + 4.0 * torch.linspace(0, 1, s + 1))[None].expand(n, s + 1).contiguous()
o, d, pa, bins = o.cuda(), d.cuda(), pa.cuda(), bins.cuda()
g_sigma = torch.randn(n, s, device="cuda") * 0.01
g_feat = torch.randn(n, s, 16, device="cuda") * 0.01
dy = torch.empty(_lib.lib().rsn_field_dy_stash_bytes(n * s), dtype=torch.uint8, device="cuda")
blob = torch.zeros(ops.wgrad_layout()[2], device="cuda")
P = n * s


def timeit(fn, flop, nbytes, name):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:28s} {ms:8.3f} ms  {flop / ms / 1e9:7.1f} TFLOP/s ({flop / ms / 1e9 / 1611.1 * 100:4.1f}% of 1611.1)  "
          f"{nbytes / ms / 1e6:7.1f} GB/s ({nbytes / ms / 1e6 / 6543.7 * 100:4.1f}% of 6543.7)")


out = {}


def fwd():
    out["f"] = ops.field_forward_train(wblob, bias, 0, o, d, pa, bins)


timeit(lambda: ops.field_forward(wblob, bias, o, d, pa, bins), P * 1230592, P * 68, "field_fwd (inference)")
timeit(fwd, P * 1230592, P * (68 + 32 + 37 * 128 + 288), "field_fwd (train, stash)")
sigma, feat, stash, aux = out["f"]
timeit(lambda: ops.field_normals(wblob_t, wd, stash, n, s), P * 1019392, P * (288 + 4 * 128 + 12), "field_chain<normals>")
timeit(lambda: ops.field_backward(wblob_t, stash, 0, o, d, pa, bins, n, s, g_sigma, g_feat, feat, aux, dy, False),
       P * 1179904, P * (288 + 35 * 128 + 160), "field_chain<backward>")
timeit(lambda: ops.field_backward(wblob_t, stash, 0, o, d, pa, bins, n, s, g_sigma, g_feat, feat, aux, dy, True),
       P * 1229056, P * (288 + 4 * 128 + 35 * 128 + 164), "field_chain<backward+area>")
timeit(lambda: ops.field_wgrad(stash, dy, P, blob), P * 1230592, P * 72 * 128, "field_wgrad")
