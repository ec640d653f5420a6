"""One launch of each field kernel at C2 primary-pass size, for `ncu --set full` (a warm-up launch of each first).
usage: ncu --set full --clock-control none --import-source on -k regex:field_ -o gpurun_out/x python scripts/ncu_kernels.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflect_sampling_nerf_b200 import _lib, ops, packing  # noqa: E402
from reflect_sampling_nerf_b200.plugin_field_init import random_field_state  # noqa: E402

n, s = 16384, 128
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
for _ in range(2):
    ops.field_forward(wblob, bias, o, d, pa, bins)
    sigma, feat, stash, aux = ops.field_forward_train(wblob, bias, 0, o, d, pa, bins)
    ops.field_normals(wblob_t, wd, stash, n, s)
    ops.field_backward(wblob_t, stash, 0, o, d, pa, bins, n, s, g_sigma, g_feat, feat, aux, dy, False)
    ops.field_wgrad(stash, dy, n * s, blob)
    torch.cuda.synchronize()
print("done")
