"""Builds csrc/*.cu into the in-tree C-ABI libraries (sm_100a only).

    python -m reflect_sampling_nerf_b200.build [--force] [--verbose]

  librsn_b200.so      the product: include/rsn_b200.h, no environment switches, no probes
  librsn_b200_dbg.so  the test build: the same sources with -DRSN_DEBUG_SWITCHES (alternative kernel forms and timing
                      ablations selected by RSN_* environment variables) + csrc/probe.cu (tcgen05 building-block
                      probes); include/rsn_b200_test.h.  Loaded only by tests/ and scripts/.

nvcc cross-compiles without a GPU.  The .so files are git-ignored but travel to the GPU box with gpurun.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librsn_b200.so")
LIB_DBG = os.path.join(HERE, "librsn_b200_dbg.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]
TEST_ONLY = {"probe.cu"}     # sources of the test build only


def _stale(lib: str) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    srcs = glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale(LIB) and not _stale(LIB_DBG):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    procs, objs = [], {LIB: [], LIB_DBG: []}
    for lib, sub, extra in ((LIB, "build", []), (LIB_DBG, os.path.join("build", "dbg"), ["-DRSN_DEBUG_SWITCHES"])):
        os.makedirs(os.path.join(HERE, sub), exist_ok=True)
        for s in srcs:  # one nvcc per translation unit, all in parallel
            if lib == LIB and os.path.basename(s) in TEST_ONLY:
                continue
            o = os.path.join(HERE, sub, os.path.basename(s)[:-3] + ".o")
            objs[lib].append(o)
            cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", s, "-o", o] + (["-Xptxas", "-v"] if verbose else [])
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    for lib in (LIB, LIB_DBG):
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib, *objs[lib]])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
