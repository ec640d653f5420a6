"""`ns-train reflect-sampling-nerf` entry point of the drop-in (reflect_sampling_nerf_config.py:27-63).

pyproject.toml registers `reflect_sampling_nerf_b200.config:reflect_sampling_nerf` under the
`nerfstudio.method_configs` entry-point group with the SAME method name, so installing this package next to (or
instead of) the reference makes `ns-train reflect-sampling-nerf --data PATH` run the B200 kernels.  Trainer
values are the reference's (100k iterations, rays/batch 1024, eval chunk 1024, RAdam 1e-3 -> 1e-4 @ 50k, Blender
dataparser, viewer).  `mixed_precision` stays True as in the reference (reflect_sampling_nerf_config.py:33): the
autocast context it opens changes nothing for the kernels (they take and return fp32 and run the MLP in bf16 with fp32
master weights and accumulation), and the GradScaler it enables multiplies the loss by a power of two that the hand-written
backward carries through exactly and the Trainer un-scales before the optimizer step.  The optimizer is the fused RAdam
(`FusedRAdamOptimizerConfig`, optim.py) under the reference's ExponentialDecay scheduler.  Only importable with
nerfstudio present (tests/test_plugin_cpu.py imports it on a stand-in nerfstudio tree).
"""
from __future__ import annotations

try:  # pragma: no cover - nerfstudio is not installed in the build image
    from nerfstudio.configs.base_config import ViewerConfig
    from nerfstudio.data.dataparsers.blender_dataparser import BlenderDataParserConfig
    from dataclasses import dataclass, field as _field
    from typing import Type

    from nerfstudio.engine.optimizers import OptimizerConfig, RAdamOptimizerConfig
    from nerfstudio.engine.schedulers import ExponentialDecaySchedulerConfig
    from nerfstudio.engine.trainer import TrainerConfig
    from nerfstudio.plugins.types import MethodSpecification

    from .model import ReflectSamplingNeRFModelConfig
    from .pipeline import ReflectSamplingNeRFDataManagerConfig, ReflectSamplingNeRFPipelineConfig

    @dataclass
    class FusedRAdamOptimizerConfig(RAdamOptimizerConfig):
        """RAdamOptimizerConfig whose optimizer is the one-launch fused RAdam (optim.py).  Optimizers.__init__ calls
        `config.setup(params=...)`; the field the parameters belong to is found through the registry the model fills."""
        _target: Type = _field(default_factory=lambda: _fused_radam_factory)

    def _fused_radam_factory(params, **kwargs):
        from .optim import FusedRAdam
        params = list(params)
        kwargs.pop("max_norm", None)
        return FusedRAdam(params, field=ReflectSamplingNeRFNerfField.owner_of(params), **kwargs)

    from .field import ReflectSamplingNeRFNerfField

    _RAYS = 1 << 10
    reflect_sampling_nerf = MethodSpecification(
        config=TrainerConfig(
            method_name="reflect-sampling-nerf",
            steps_per_eval_batch=100, steps_per_save=1000, max_num_iterations=100000, mixed_precision=True,
            pipeline=ReflectSamplingNeRFPipelineConfig(
                datamanager=ReflectSamplingNeRFDataManagerConfig(
                    dataparser=BlenderDataParserConfig(), train_num_rays_per_batch=_RAYS,
                    eval_num_rays_per_batch=_RAYS),
                model=ReflectSamplingNeRFModelConfig(eval_num_rays_per_chunk=_RAYS)),
            optimizers={"fields": {   # the model exposes only this group (model.py:134-139, App. B Q15)
                "optimizer": FusedRAdamOptimizerConfig(lr=1e-3, eps=1e-15),
                "scheduler": ExponentialDecaySchedulerConfig(lr_final=1e-4, max_steps=50000)}},
            viewer=ViewerConfig(num_rays_per_chunk=_RAYS), vis="viewer"),
        description="reflect-sampling-nerf on hand-written sm_100a (B200) kernels.")
except ImportError as exc:  # noqa: F841
    reflect_sampling_nerf = None
