"""Pipeline boundary of the drop-in (reflect_sampling_nerf_pipeline.py:26-91).

Two things live in the reference pipeline that touch the hot path:
  * the loss-coefficient warm-up of get_train_loss_dict (pipeline.py:79-91): normal / orientation terms are off for
    step < 50 -- `warmup_loss_coefficients` below, nerfstudio-free so it is testable anywhere;
  * the data-parallel wrapper (pipeline.py:73-77): DistributedDataParallel(find_unused_parameters=True) + barrier.
    Here the model is NOT wrapped: the wgrad kernels write one flat gradient blob and train_path._flush_grads
    all-reduces it once per step (`field.dp_world_size = world_size`).  What the wrapper also did at construction --
    broadcast rank 0's parameters, nerfstudio seeds every rank differently -- is `train_path.sync_parameters`.

The nerfstudio-facing classes exist only when nerfstudio is importable (it is not in the build image).
"""
from __future__ import annotations

from typing import Dict

WARMUP_STEPS = 50
_WARM = {"predicted_normal_loss_coarse": 3e-5, "predicted_normal_loss_fine": 3e-4,
         "orientation_loss_coarse": 1e-2, "orientation_loss_fine": 1e-1}


def warmup_loss_coefficients(step: int, coefficients: Dict[str, float]) -> Dict[str, float]:
    """pipeline.py:79-91 (mutates and returns the model's coefficient dict, as the reference does; App. B Q14)."""
    for key, value in _WARM.items():
        coefficients[key] = value if step >= WARMUP_STEPS else 0.0
    return coefficients


try:  # pragma: no cover - nerfstudio is not installed in the build image
    import typing
    from dataclasses import dataclass, field
    from typing import Literal, Optional, Type

    import torch.distributed as dist
    from nerfstudio.data.datamanagers.base_datamanager import (DataManagerConfig, VanillaDataManager,
                                                               VanillaDataManagerConfig)
    from nerfstudio.models.base_model import ModelConfig
    from nerfstudio.pipelines.base_pipeline import VanillaPipeline, VanillaPipelineConfig

    from .model import ReflectSamplingNeRFModelConfig

    @dataclass
    class ReflectSamplingNeRFDataManagerConfig(VanillaDataManagerConfig):
        """reflect_sampling_nerf_datamanager.py:17-24 (the datamanager is out of the hot path and unchanged)."""
        _target: Type = field(default_factory=lambda: VanillaDataManager)

    @dataclass
    class ReflectSamplingNeRFPipelineConfig(VanillaPipelineConfig):
        _target: Type = field(default_factory=lambda: ReflectSamplingNeRFPipeline)
        datamanager: DataManagerConfig = field(default_factory=ReflectSamplingNeRFDataManagerConfig)
        model: ModelConfig = field(default_factory=ReflectSamplingNeRFModelConfig)

    class ReflectSamplingNeRFPipeline(VanillaPipeline):
        def __init__(self, config, device: str, test_mode: Literal["test", "val", "inference"] = "val",
                     world_size: int = 1, local_rank: int = 0, grad_scaler=None):
            super(VanillaPipeline, self).__init__()
            self.config, self.test_mode = config, test_mode
            self.datamanager = config.datamanager.setup(device=device, test_mode=test_mode, world_size=world_size,
                                                        local_rank=local_rank)
            self.datamanager.to(device)
            assert self.datamanager.train_dataset is not None, "Missing input dataset"
            self._model = config.model.setup(scene_box=self.datamanager.train_dataset.scene_box,
                                             num_train_data=len(self.datamanager.train_dataset),
                                             metadata=self.datamanager.train_dataset.metadata, device=device,
                                             grad_scaler=grad_scaler)
            self.model.to(device)
            self.world_size = world_size
            self._model.field.dp_world_size = world_size      # flat-gradient all-reduce instead of the DDP wrapper
            if world_size > 1:
                from .train_path import sync_parameters
                sync_parameters(self._model.field)            # DDP's construction-time broadcast (pipeline.py:73-77)
                dist.barrier(device_ids=[local_rank])

        def get_train_loss_dict(self, step: int):
            warmup_loss_coefficients(step, self.model.config.loss_coefficients)
            return super().get_train_loss_dict(step)

        def get_eval_image_metrics_and_images(self, step: int):
            """The reference inherits VanillaPipeline's, which dies on its model's KeyError (App. B Q13); the fixed model
            method makes the inherited flow work unchanged."""
            return super().get_eval_image_metrics_and_images(step)

    HAVE_NERFSTUDIO = True
except Exception:  # noqa: BLE001
    HAVE_NERFSTUDIO = False
