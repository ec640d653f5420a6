"""Shared test helpers (synthetic rays, golden loading, oracle construction)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED_WEIGHTS = 20261018           # tests/golden/make_golden.py
GOLDEN_SIZES = dict(num_coarse_samples=24, num_importance_samples=24,
                    num_reflect_coarse_samples=12, num_reflect_importance_samples=12)


def load_golden(mode: str):
    z = np.load(os.path.join(GOLDEN_DIR, f"refpath_{mode}.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def oracle_model(sizes=None, seed=SEED_WEIGHTS):
    from oracle.refpath import OracleModel
    torch.manual_seed(seed)
    return OracleModel(**(sizes or GOLDEN_SIZES))


def synthetic_rays(n: int, seed: int, pixel_area: float = 3.2e-6, device="cpu"):
    """directions ~ normalised N(0,I), origins = -4 d + 0.3 N(0,I)."""
    g = torch.Generator().manual_seed(seed)
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    o = -4.0 * d + 0.3 * torch.randn(n, 3, generator=g)
    area = torch.full((n, 1), pixel_area)
    image = torch.rand(n, 3, generator=g)
    return tuple(t.to(device) for t in (o, d, area, image))
