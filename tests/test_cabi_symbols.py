"""CPU-only: the C-ABI libraries load and export every symbol their headers declare -- the product library
(include/rsn_b200.h) and the test build (include/rsn_b200_test.h) -- and the product library contains none of the
test-only entry points or environment reads."""
import ctypes
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    text = open(os.path.join(REPO, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from reflect_sampling_nerf_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols("rsn_b200.h")
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    # and the loader's signature table covers the header exactly
    assert sorted(_lib._SIGNATURES) == syms
    assert lib.rsn_version() >= 200


def test_test_build_exports_both_headers_and_product_has_no_test_symbols():
    import __graft_entry__
    __graft_entry__.build()
    from reflect_sampling_nerf_b200 import _lib
    dbg = ctypes.CDLL(_lib.LIB_DBG_PATH)
    prod = ctypes.CDLL(_lib.LIB_PATH)
    test_only = declared_symbols("rsn_b200_test.h")
    assert sorted(_lib._SIGNATURES_DBG) == test_only
    for s in declared_symbols("rsn_b200.h") + test_only:
        assert hasattr(dbg, s), s
    for s in test_only:
        assert not hasattr(prod, s), f"{s} must not ship in the product library"
    # our translation units of the product build never read the environment (the CUDA runtime linked into the .so does,
    # for its own CUDA_* variables); those of the test build do
    bdir = os.path.join(REPO, "reflect_sampling_nerf_b200", "build")

    def objects_importing_getenv(d):
        hits = []
        for f in sorted(os.listdir(d)):
            if f.endswith(".o"):
                out = subprocess.run(["nm", "--undefined-only", os.path.join(d, f)], capture_output=True, text=True).stdout
                if "getenv" in out:
                    hits.append(f)
        return hits
    assert objects_importing_getenv(bdir) == []
    assert objects_importing_getenv(os.path.join(bdir, "dbg"))


def test_ops_fail_loudly_without_cuda():
    import pytest
    import torch
    from reflect_sampling_nerf_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.sample_spaced(torch.zeros(4, 1), torch.ones(4, 1), 8, ops.UNIFORM)


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import pytest
    from reflect_sampling_nerf_b200 import _lib
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib._load(str(tmp_path / "librsn_b200.so"), _lib._SIGNATURES)
