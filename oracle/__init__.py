"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU/PyTorch-fp32 restatement of the per-ray rendering hot path of
236088/reflect-sampling-nerf (reference files reflect_sampling_nerf_model.py,
_field.py, _components.py) and of the un-vendored upstream nerfstudio primitives
those files call.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import anything from this package, and only as the
checker / the reported CPU baseline.  The shipped package
`reflect_sampling_nerf_b200` never imports it.

PARITY PINNING STATUS (see DESIGN.md "Oracle"):
  * `oracle.refpath` (the restatement of the three reference files) is PINNED:
    `tests/golden/make_golden.py` executes the UNMODIFIED reference files from
    /root/reference on top of `oracle.nerfstudio_shim` (stand-in `nerfstudio`
    module tree backed by `oracle.upstream`) and commits the resulting vectors
    under tests/golden/; tests/test_oracle_golden.py checks refpath against them.
  * `oracle.upstream` (nerfstudio samplers / renderers / encodings / MLP / heads)
    is UNPINNED: nerfstudio is a third-party dependency (`nerfstudio >= 0.3.0`,
    no lock file, pyproject.toml:6) absent from /root/reference and from this
    image; it is restated from its published 0.3.x/1.0.x behaviour.  The
    reference repo ships no tests, fixtures or golden vectors.
"""
