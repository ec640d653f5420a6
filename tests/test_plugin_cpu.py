"""CPU: the nerfstudio-facing half of the drop-in (`ns-train reflect-sampling-nerf` entry point,
reflect_sampling_nerf_config.py:27-63 / reflect_sampling_nerf_pipeline.py:26-91) imported and instantiated on a stand-in
nerfstudio tree (tests/nerfstudio_standin.py) in a fresh interpreter -- the real nerfstudio is not in this image."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, sys
sys.path.insert(0, {repo!r}); sys.path.insert(0, {tests!r})
import nerfstudio_standin as S
S.install()
import torch
from reflect_sampling_nerf_b200 import config, model, pipeline, rays
from reflect_sampling_nerf_b200.optim import FusedRAdam
spec = config.reflect_sampling_nerf
out = dict(have=[model.HAVE_NERFSTUDIO, pipeline.HAVE_NERFSTUDIO, rays.HAVE_NERFSTUDIO],
           spec=type(spec).__name__, method=spec.config.method_name, mixed=spec.config.mixed_precision,
           iters=spec.config.max_num_iterations, rays=spec.config.pipeline.datamanager.train_num_rays_per_batch,
           chunk=spec.config.pipeline.model.eval_num_rays_per_chunk, groups=sorted(spec.config.optimizers),
           lr=spec.config.optimizers["fields"]["optimizer"].lr, eps=spec.config.optimizers["fields"]["optimizer"].eps,
           lr_final=spec.config.optimizers["fields"]["scheduler"].lr_final,
           sched_steps=spec.config.optimizers["fields"]["scheduler"].max_steps)
# Trainer.setup -> PipelineConfig.setup -> ModelConfig.setup (reflect_sampling_nerf_pipeline.py:52-77), single process on CPU
pipe = spec.config.pipeline.setup(device="cpu", test_mode="val", world_size=1, local_rank=0, grad_scaler=None)
m = pipe.model
out["pipe"] = type(pipe).__name__
out["model_bases"] = [c.__module__ for c in type(m).__mro__[1:3]]
out["collider"] = type(m.collider).__module__
out["param_groups"] = sorted(m.get_param_groups())
out["n_params"] = sum(p.numel() for p in m.get_param_groups()["fields"])
out["attrs"] = all(hasattr(m, a) for a in ("field", "sampler_uniform", "sampler_pdf", "sampler_reciprocal", "sampler_reflect_pdf",
    "renderer_rgb", "renderer_accumulation", "renderer_depth", "renderer_normals", "renderer_roughness", "renderer_factor",
    "renderer_reflect", "rgb_loss", "psnr", "ssim", "near", "far"))
out["dp_world"] = m.field.dp_world_size
# Optimizers(config.optimizers, param_groups): config.setup(params=...)
opt = spec.config.optimizers["fields"]["optimizer"].setup(params=m.get_param_groups()["fields"])
out["opt"] = [type(opt).__name__, isinstance(opt, FusedRAdam), isinstance(opt, torch.optim.Optimizer), opt.field is m.field,
              opt.param_groups[0]["lr"], opt.param_groups[0]["eps"]]
# the warm-up rewrite of pipeline.py:79-91
pipeline.warmup_loss_coefficients(10, m.config.loss_coefficients)
out["warm10"] = m.config.loss_coefficients["orientation_loss_fine"]
pipeline.warmup_loss_coefficients(50, m.config.loss_coefficients)
out["warm50"] = m.config.loss_coefficients["orientation_loss_fine"]
print("RESULT " + json.dumps(out))
"""


def test_method_specification_builds_on_a_stand_in_nerfstudio():
    code = SCRIPT.format(repo=REPO, tests=os.path.join(REPO, "tests"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    r = json.loads(line[len("RESULT "):])
    assert r["have"] == [True, True, True]
    assert r["spec"] == "MethodSpecification" and r["method"] == "reflect-sampling-nerf"
    # reference values (reflect_sampling_nerf_config.py:27-63)
    assert r["mixed"] is True and r["iters"] == 100000 and r["rays"] == 1024 and r["chunk"] == 1024
    assert r["groups"] == ["fields"] and r["lr"] == 1e-3 and r["eps"] == 1e-15
    assert r["lr_final"] == 1e-4 and r["sched_steps"] == 50000
    assert r["pipe"] == "ReflectSamplingNeRFPipeline"
    assert r["model_bases"][0] == "nerfstudio_standin" or r["model_bases"][0].startswith("nerfstudio")
    assert r["collider"].startswith("nerfstudio")          # the base Model's collider, not the fallback's
    assert r["param_groups"] == ["fields"] and r["n_params"] == 618513 and r["attrs"] and r["dp_world"] == 1
    assert r["opt"] == ["FusedRAdam", True, True, True, 1e-3, 1e-15]
    assert r["warm10"] == 0.0 and r["warm50"] == 1e-1


def test_fallback_classes_without_nerfstudio():
    from reflect_sampling_nerf_b200 import config, model, pipeline
    assert model.HAVE_NERFSTUDIO is False and pipeline.HAVE_NERFSTUDIO is False
    assert config.reflect_sampling_nerf is None
    m = model.ReflectSamplingNeRFModel(model.ReflectSamplingNeRFModelConfig())
    assert isinstance(m.collider, model._NearFarCollider)
    assert sum(p.numel() for p in m.get_param_groups()["fields"]) == 618513
