"""Sanity + timing of the other BASELINE.json configs through the public model API (render-only ones):
  C3: full 800x800 frame = 640,000 rays, eval mode, chunks of 65,536 rays, 128+128 / 64+64 samples
  C5: one GPU's shard of the stress config = 8,192 rays x (256 + 256) samples + (128 + 128) reflected
usage: python scripts/bench_configs.py"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from reflect_sampling_nerf_b200.model import ReflectSamplingNeRFModel, ReflectSamplingNeRFModelConfig
from reflect_sampling_nerf_b200.rays import RayBundle


def run(name, n_rays, chunk, cfg, area):
    torch.manual_seed(0)
    model = ReflectSamplingNeRFModel(ReflectSamplingNeRFModelConfig(**cfg)).cuda().eval()
    o, d, a, _ = [t.cuda() for t in bench.synthetic_batch(n_rays, 7)]
    a = torch.full_like(a, area)

    def frame():
        outs = []
        for i in range(0, n_rays, chunk):
            out = model(RayBundle(origins=o[i:i + chunk], directions=d[i:i + chunk], pixel_area=a[i:i + chunk]))
            outs.append((out["mid_rgb_fine"], out["mid_reflect_fine"], out["mask"]))
        return outs
    with torch.no_grad():
        frame()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        outs = frame()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    rgb = torch.cat([x[0] for x in outs]); refl = torch.cat([x[1] for x in outs]); mask = torch.cat([x[2] for x in outs])
    assert rgb.shape == (n_rays, 3) and bool(torch.isfinite(rgb).all()) and bool(torch.isfinite(refl).all())
    assert float(rgb.min()) >= 0 and float(rgb.max()) <= 1
    print(f"{name}: {n_rays} rays in {dt*1e3:.1f} ms = {n_rays/dt/1e6:.2f} M rays/s; masked {int(mask.sum())}")


run("C3 800x800 frame (eval, 128+128 / 64+64)", 640000, 65536, bench.CFG, 8.1e-7)
run("C5 shard 8192 rays (256+256 / 128+128)", 8192, 8192,
    dict(num_coarse_samples=256, num_importance_samples=256, num_reflect_coarse_samples=128,
         num_reflect_importance_samples=128), 3.2e-6)
run("reference default 1024 rays (128+128 / 64+64)", 1024, 1024, bench.CFG, 3.2e-6)
