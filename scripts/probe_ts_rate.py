"""Issue rate of tcgen05.mma with the A operand in TMEM vs shared memory (cycles per M128 x N x K16 MMA)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflect_sampling_nerf_b200 import _lib  # noqa: E402
from reflect_sampling_nerf_b200.blocks import pack_blocks  # noqa: E402

iters = 4096
for N in (256, 128, 64):
    x = torch.randn(128, 256).bfloat16()
    w = torch.randn(N, 256).bfloat16()
    xb, wb = pack_blocks(x).cuda(), pack_blocks(w).cuda()
    out = torch.empty(128, N, device="cuda")
    cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
    _lib.call("rsn_probe_umma_ts", xb.data_ptr(), wb.data_ptr(), N, 4, out.data_ptr(), iters, cyc.data_ptr(), _lib.stream())
    torch.cuda.synchronize()
    ts = cyc.item() / iters
    _lib.call("rsn_probe_umma_rate", 2, 0, N, iters, cyc.data_ptr(), _lib.stream())
    torch.cuda.synchronize()
    ss = cyc.item() / iters
    print(f"N={N}: A in TMEM {ts:.1f} cycles/MMA, A in shared memory {ss:.1f} cycles/MMA")
