"""Generate golden vectors by executing the UNMODIFIED reference files.

    python tests/golden/make_golden.py        # needs /root/reference (this container only)

The three hot-path files under /root/reference/reflect_sampling_nerf are imported as they are,
on top of `oracle.nerfstudio_shim` (stand-in module tree for the un-vendored nerfstudio /
nerfacc / torchmetrics imports).  Their outputs, losses and a digest of the gradients are written
to tests/golden/refpath_{train,eval}.npz.  Weights are the seeded default nn.Linear init
(a checksum is stored so a changed RNG stream is detected rather than silently compared).

What this pins: the reference's own code (model.py / field.py / components.py).
What it does not pin: nerfstudio itself (restated in oracle/upstream.py) -- see oracle/__init__.py.
"""
import contextlib
import io
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")   # `reflect_sampling_nerf` must resolve to the reference
sys.path.insert(1, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import nerfstudio_shim, upstream as U  # noqa: E402

SEED_WEIGHTS = 20261018
SIZES = dict(num_coarse_samples=24, num_importance_samples=24,
             num_reflect_coarse_samples=12, num_reflect_importance_samples=12)
N_RAYS = 96


def synthetic_rays(n: int, seed: int):
    """Cameras on a radius-4 sphere looking roughly at the origin."""
    g = torch.Generator().manual_seed(seed)
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    o = -4.0 * d + 0.3 * torch.randn(n, 3, generator=g)
    area = torch.full((n, 1), 3.2e-6)
    image = torch.rand(n, 3, generator=g)
    return o, d, area, image


def jitters(n: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    return dict(
        uniform=torch.rand(n, SIZES["num_coarse_samples"] + 1, generator=g),
        pdf=torch.rand(n, SIZES["num_importance_samples"] + 1, generator=g),
        # reflected passes run on M <= n rays; rows are taken from the top
        reciprocal=torch.rand(n, SIZES["num_reflect_coarse_samples"] + 1, generator=g),
        reflect_pdf=torch.rand(n, SIZES["num_reflect_importance_samples"] + 1, generator=g),
    )


def weight_checksum(model) -> float:
    return float(sum(p.detach().double().abs().sum() for p in model.field.parameters()))


def main() -> None:
    nerfstudio_shim.install()
    from reflect_sampling_nerf.reflect_sampling_nerf_model import ReflectSamplingNeRFModelConfig
    import reflect_sampling_nerf
    assert reflect_sampling_nerf.__path__[0].startswith("/root/reference"), reflect_sampling_nerf.__path__

    for mode in ("train", "eval"):
        torch.manual_seed(SEED_WEIGHTS)
        model = ReflectSamplingNeRFModelConfig(**SIZES).setup(scene_box=None, num_train_data=1)
        model.train(mode == "train")
        o, d, area, image = synthetic_rays(N_RAYS, seed=7)
        jit = jitters(N_RAYS, seed=11)
        bundle = U.RayBundle(origins=o, directions=d, pixel_area=area)
        bundle = model.collider(bundle)

        # inject jitter: primary samplers directly; reflected samplers need the mask size M first.
        model.sampler_uniform.injected_rand = jit["uniform"]
        model.sampler_pdf.injected_rand = jit["pdf"]
        # first run to learn M (weights do not depend on the reflected jitter)
        with contextlib.redirect_stdout(io.StringIO()), torch.no_grad() if mode == "eval" else contextlib.nullcontext():
            probe = model.get_outputs(bundle)
        m = int(probe["mask"].sum())
        model.sampler_reciprocal.injected_rand = jit["reciprocal"][:m]
        model.sampler_reflect_pdf.injected_rand = jit["reflect_pdf"][:m]
        model.zero_grad()

        with contextlib.redirect_stdout(io.StringIO()):
            if mode == "train":
                out = model.get_outputs(bundle)
                loss = model.get_loss_dict(out, {"image": image})
                sum(loss.values()).backward()
            else:
                with torch.no_grad():
                    out = model.get_outputs(bundle)
                    loss = model.get_loss_dict(out, {"image": image})

        rec = {"in_origins": o, "in_directions": d, "in_pixel_area": area, "in_image": image,
               "in_nears": bundle.nears, "in_fars": bundle.fars, "num_masked": torch.tensor(m),
               "weight_checksum": torch.tensor(weight_checksum(model), dtype=torch.float64)}
        rec.update({f"jit_{k}": v for k, v in jit.items()})
        rec.update({f"out_{k}": v.detach() for k, v in out.items()})
        rec.update({f"loss_{k}": v.detach() for k, v in loss.items()})
        if mode == "train":
            for name, p in model.field.named_parameters():
                g = p.grad if p.grad is not None else torch.zeros_like(p)
                rec[f"gsum_{name}"] = g.double().sum()
                rec[f"gabs_{name}"] = g.double().abs().sum()
                if p.numel() <= 1024 or name.endswith("bias"):
                    rec[f"grad_{name}"] = g
            # a strided slice of one big matrix per block, enough to catch a wrong wgrad layout
            rec["grad_mlp_base.layers.4.weight[::16,::16]"] = model.field.mlp_base.layers[4].weight.grad[::16, ::16]
            rec["grad_mlp_mid.layers.0.weight[::8,::8]"] = model.field.mlp_mid.layers[0].weight.grad[::8, ::8]
        path = os.path.join(HERE, f"refpath_{mode}.npz")
        np.savez_compressed(path, **{k: v.detach().cpu().numpy() for k, v in rec.items()})
        print(f"wrote {path}: M={m}, {len(rec)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
