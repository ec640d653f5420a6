"""Per-kernel SASS opcode summary of librsn_b200.so (cuobjdump -sass): counts of the mnemonics that prove a
Blackwell-native kernel (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk,
UTMALDG/UTMASTG = tensor-map TMA, SYNCS = mbarrier, HMMA = legacy mma.sync (must be 0).
    python scripts/sass_summary.py [lib.so] > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "reflect_sampling_nerf_b200", "librsn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "UTCBAR", "REDG", "RED"]
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        kernels[cur]["_total"] += 1
        if op in OPS:
            kernels[cur][op] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.basename(lib)}: {len(kernels)} kernels, arch sm_100a (cuobjdump -sass)")
print(f"{'kernel':70s} {'instr':>7s} " + " ".join(f"{o:>8s}" for o in OPS[:9]))
tot = collections.Counter()
for (name, c), dn in zip(kernels.items(), demangle):
    short = re.sub(r"\(anonymous namespace\)::", "", dn)
    short = re.sub(r"\(.*", "", short)[:70]
    print(f"{short:70s} {c['_total']:7d} " + " ".join(f"{c[o]:8d}" for o in OPS[:9]))
    tot.update(c)
print(f"{'TOTAL':70s} {tot['_total']:7d} " + " ".join(f"{tot[o]:8d}" for o in OPS[:9]))
