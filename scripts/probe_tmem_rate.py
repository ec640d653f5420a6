"""TMEM read/write throughput of epilogue-style warps (tcgen05.ld / tcgen05.st), alone and under a concurrent
tcgen05.mma stream; cycles per 64-column x 32-lane slice."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflect_sampling_nerf_b200 import _lib  # noqa: E402

iters = 1024
names = {0: "2 x ld.x32 + wait", 1: "4 x ld.x16 + wait", 2: "2 x ld.x32 + wait + st.x32 + wait", 3: "2 x ld.x32, wait every 4th"}
for mma in (0, 1):
    for mode in (0, 1, 2, 3):
        for nw in (1, 4, 8):
            cyc = torch.zeros(9, dtype=torch.int64, device="cuda")
            # the MMA stream is sized to outlast the readers (128 cycles per MMA)
            mma_iters = 0 if not mma else 4 * ((iters * 1200 // 128) // 4)
            _lib.call("rsn_probe_tmem_rate", nw, mode, iters, mma_iters, cyc.data_ptr(), _lib.stream())
            torch.cuda.synchronize()
            c = cyc[:nw].max().item() / iters
            extra = f"   MMA {cyc[8].item() / mma_iters:6.1f} cycles each" if mma else ""
            print(f"{'MMA busy' if mma else 'MMA idle'} {names[mode]:36s} warps={nw}: {c:7.1f} cycles/iter  -> "
                  f"{nw * 8192 / c:7.1f} B/cycle/SM{extra}")
