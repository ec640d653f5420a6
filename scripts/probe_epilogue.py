"""Step-by-step cost of one epilogue group (64 accumulator columns x 128 rows on four warps), alone and under a
concurrent tcgen05.mma stream."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflect_sampling_nerf_b200 import _lib  # noqa: E402

iters = 1024
cases = [(0, "tcgen05.ld x2 + wait"), (1, "+ bias(reg) relu pack"), (17, "+ bias from constant memory"),
         (19, "+ st.shared x8"), (23, "+ fence.proxy.async"), (31, "+ tc fence + arrive (128 threads)"),
         (95, "  ... arrive once per warp"), (49, "TS: math + tcgen05.st + wait::st"), (57, "TS: + tc fence + arrive"),
         (121, "TS:   ... arrive once per warp")]
for mma in (0, 1):
    for steps, name in cases:
        cyc = torch.zeros(9, dtype=torch.int64, device="cuda")
        mma_iters = 0 if not mma else 4 * ((iters * 1500 // 128) // 4)
        _lib.call("rsn_probe_epilogue", steps, iters, mma_iters, cyc.data_ptr(), _lib.stream())
        torch.cuda.synchronize()
        c = cyc[:4].max().item() / iters
        extra = f"   MMA {cyc[8].item() / mma_iters:6.1f} cycles each" if mma else ""
        print(f"{'MMA busy' if mma else 'MMA idle'} {name:40s}: {c:7.1f} cycles/group{extra}")
