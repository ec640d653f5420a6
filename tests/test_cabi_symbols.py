"""CPU-only: the C-ABI library loads and exports every symbol include/rsn_b200.h declares."""
import ctypes
import os
import re

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(REPO, "include", "rsn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from reflect_sampling_nerf_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 8
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    # and the loader's signature table covers the header exactly
    assert sorted(_lib._SIGNATURES) == syms
    assert lib.rsn_version() >= 100


def test_ops_fail_loudly_without_cuda():
    import pytest
    import torch
    from reflect_sampling_nerf_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.sample_spaced(torch.zeros(4, 1), torch.ones(4, 1), 8, ops.UNIFORM)
