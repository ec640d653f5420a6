"""In-kernel cycle trace of the forward field kernel (test build, RSN_FWD_DEBUG=32): where one tile's layer period goes.
usage: python scripts/trace_fwd.py [train|infer]   (prints per-layer timelines of the third tile of CTA 0, in SM cycles)"""
import ctypes
import os
import sys

os.environ["RSN_FWD_DEBUG"] = "32"
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflect_sampling_nerf_b200 import _lib, ops, packing  # noqa: E402
from reflect_sampling_nerf_b200.plugin_field_init import random_field_state  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
_lib.use_dbg(True)
n, s = 16384, 128
torch.manual_seed(0)
sd = random_field_state()
wblob, bias = [t.cuda() for t in packing.pack_field(sd)]
d = torch.nn.functional.normalize(torch.randn(n, 3), dim=-1)
o = -4 * d + 0.3 * torch.randn(n, 3)
pa = torch.full((n,), 3.2e-6)
bins = (2.0 + 4.0 * torch.linspace(0, 1, s + 1))[None].expand(n, s + 1).contiguous()
o, d, pa, bins = o.cuda(), d.cuda(), pa.cuda(), bins.cuda()
buf = (ctypes.c_longlong * 8192)()
for rep in range(2):
    if mode == "train":
        ops.field_forward_train(wblob, bias, 0, o, d, pa, bins)
    else:
        ops.field_forward(wblob, bias, o, d, pa, bins)
    npairs = _lib.lib_dbg().rsn_debug_fwd_trace(buf, 4095)
ev = sorted(((buf[2 * i + 1], buf[2 * i]) for i in range(npairs)))
t0 = ev[0][0]
print(f"# {mode}: {npairs} events; cycles relative to the first event")
names = {10: "issuer act_ready", 11: "issuer ring_ready", 12: "issuer mma issued", 13: "issuer ring released", 15: "issuer commit",
         20: "epi acc_full", 21: "epi handover", 22: "epi staged", 23: "epi  stage: guard done", 24: "epi  stage: row in smem"}
for t, tag in ev:
    kind = tag // 100
    l, g = (tag % 100) // 10, tag % 10
    if kind in (15, 20):
        l, g = tag % 100, -1
    print(f"{t - t0:8d}  {names.get(kind, kind):18s} layer {l} group {g}")
# summary: per-layer period seen by the epilogue, group conversion times
acc = {tag % 100: t for t, tag in ev if tag // 100 == 20}
hand = {((tag % 100) // 10, tag % 10): t for t, tag in ev if tag // 100 == 21}
com = {tag % 100: t for t, tag in ev if tag // 100 == 15}
print("# layer: acc_full->g0, g0->g1, g1->g2, g2->g3 | commit(l+1) - handover(l, g3) | period (acc_full l+1 - acc_full l)")
for l in range(8):
    if l in acc and (l, 3) in hand:
        gs = [hand[(l, 0)] - acc[l]] + [hand[(l, g)] - hand[(l, g - 1)] for g in (1, 2, 3)]
        tail = com.get(l + 1, 0) - hand[(l, 3)] if (l + 1) in com else None
        per = acc[l + 1] - acc[l] if (l + 1) in acc else None
        print(f"  {l}: {gs} | {tail} | {per}")
