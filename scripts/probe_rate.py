"""tcgen05.mma issue-rate probe (results quoted in DESIGN.md)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflect_sampling_nerf_b200 import _lib
out = torch.zeros(1, dtype=torch.int64, device="cuda")
def run(a, b, n, cheap=0, grid=1, iters=4096):
    _lib.call("rsn_probe_umma_rate", a | (cheap << 1) | (grid << 8), b, n, iters, out.data_ptr(), _lib.stream())
    torch.cuda.synchronize()
    return out.item() / iters
for cheap in (0, 1):
    for n in (256, 128, 64, 16):
        r = [run(a, b, n, cheap) for a, b in ((0, 0), (1, 0), (0, 1), (1, 1))]
        print(f"issue={'lo-add ' if cheap else 'rebuild'} N={n:3d}  K/K {r[0]:.1f}  MN/K {r[1]:.1f}  K/MN {r[2]:.1f}  MN/MN {r[3]:.1f} cycles/MMA")
print(f"148 CTAs, K/K N=256 lo-add: {run(0, 0, 256, 1, 148):.1f}")
def run2(n, commit, switch, iters=4096):
    _lib.call("rsn_probe_umma_rate", 0 | (1 << 1) | (1 << 8) | (commit << 16) | (switch << 20), 0, n, iters, out.data_ptr(), _lib.stream())
    torch.cuda.synchronize()
    return out.item() / iters
for n in (256, 128):
    for commit in (0, 1, 2, 4):
        for switch in (0, 1, 2, 4):
            print(f"N={n} commit every {4*commit:2d} MMAs, switch accumulator every {4*switch:2d} MMAs: {run2(n, commit, switch):.1f} cycles/MMA")
for n in (256, 128, 16):
    for pairs in (1, 74):
        _lib.call("rsn_probe_umma_rate_2cta", n, 4096, pairs, out.data_ptr(), _lib.stream())
        torch.cuda.synchronize()
        print(f"cta_group::2 M=256 N={n} on {pairs} pair(s): {out.item() / 4096:.1f} cycles/MMA")
