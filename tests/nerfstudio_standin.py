"""TEST INFRASTRUCTURE: a stand-in `nerfstudio` module tree, just deep enough to import and instantiate the plugin's
nerfstudio-facing half (reflect_sampling_nerf_b200/config.py, pipeline.py, and model.py / rays.py on their nerfstudio
branches) in an interpreter that has no nerfstudio (tests/test_plugin_cpu.py runs it in a subprocess).  The classes mirror
the constructor / dataclass contracts the reference relies on, not upstream behaviour.
`install()` refuses to shadow a real nerfstudio."""
from __future__ import annotations

import importlib.util
import sys
import types
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, Optional, Type

import torch
from torch import nn


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], child, m)
    return m


@dataclass
class InstantiateConfig:
    _target: Type = object

    def setup(self, **kwargs) -> Any:
        return self._target(self, **kwargs)


@dataclass
class ViewerConfig:
    num_rays_per_chunk: int = 32768


@dataclass
class OptimizerConfig(InstantiateConfig):
    _target: Type = torch.optim.Adam
    lr: float = 0.0005
    eps: float = 1e-08
    max_norm: Optional[float] = None

    def setup(self, params) -> torch.optim.Optimizer:   # upstream engine/optimizers.py
        kwargs = vars(self).copy()
        kwargs.pop("_target")
        kwargs.pop("max_norm")
        return self._target(params, **kwargs)


@dataclass
class RAdamOptimizerConfig(OptimizerConfig):
    _target: Type = torch.optim.RAdam
    weight_decay: float = 0


@dataclass
class ExponentialDecaySchedulerConfig:
    lr_pre_warmup: float = 1e-8
    lr_final: Optional[float] = None
    warmup_steps: int = 0
    max_steps: int = 100000
    ramp: str = "cosine"


@dataclass
class BlenderDataParserConfig:
    data: str = "data/blender/lego"
    scale_factor: float = 1.0
    alpha_color: str = "white"


@dataclass
class DataManagerConfig(InstantiateConfig):
    pass


class _Dataset:
    scene_box = "scene_box"
    metadata: Dict[str, Any] = {}

    def __len__(self) -> int:
        return 100


class VanillaDataManager(nn.Module):
    def __init__(self, config, device="cpu", test_mode="val", world_size=1, local_rank=0, **kwargs) -> None:
        super().__init__()
        self.config, self.device_, self.world_size, self.local_rank = config, device, world_size, local_rank
        self.train_dataset = _Dataset()
        self.train_count = 0


@dataclass
class VanillaDataManagerConfig(DataManagerConfig):
    _target: Type = field(default_factory=lambda: VanillaDataManager)
    dataparser: Any = field(default_factory=BlenderDataParserConfig)
    train_num_rays_per_batch: int = 1024
    eval_num_rays_per_batch: int = 1024


@dataclass
class ModelConfig(InstantiateConfig):
    _target: Type = object
    enable_collider: bool = True
    collider_params: Optional[Dict[str, float]] = field(default_factory=lambda: {"near_plane": 2.0, "far_plane": 6.0})
    loss_coefficients: Dict[str, float] = field(default_factory=dict)
    eval_num_rays_per_chunk: int = 4096


class NearFarCollider(nn.Module):
    def __init__(self, near_plane: float, far_plane: float, **kwargs) -> None:
        super().__init__()
        self.near_plane, self.far_plane = near_plane, far_plane

    def forward(self, ray_bundle):
        ones = torch.ones_like(ray_bundle.origins[..., 0:1])
        ray_bundle.nears = ones * (self.near_plane if self.training else 0.0)
        ray_bundle.fars = ones * self.far_plane
        return ray_bundle


class Model(nn.Module):
    def __init__(self, config, scene_box=None, num_train_data: int = 0, **kwargs) -> None:
        super().__init__()
        self.config, self.scene_box, self.num_train_data, self.kwargs = config, scene_box, num_train_data, kwargs
        self.collider = None
        self.populate_modules()
        self.device_indicator_param = nn.Parameter(torch.empty(0))

    @property
    def device(self):
        return self.device_indicator_param.device

    def populate_modules(self):
        if self.config.enable_collider:
            self.collider = NearFarCollider(near_plane=self.config.collider_params["near_plane"],
                                            far_plane=self.config.collider_params["far_plane"])

    def forward(self, ray_bundle):
        if self.collider is not None:
            ray_bundle = self.collider(ray_bundle)
        return self.get_outputs(ray_bundle)

    def get_metrics_dict(self, outputs, batch):
        return {}


class Pipeline(nn.Module):
    @property
    def model(self):
        return self._model

    @property
    def device(self):
        return self.model.device


class VanillaPipeline(Pipeline):
    def get_train_loss_dict(self, step: int):
        ray_bundle, batch = self.datamanager.next_train(step)
        model_outputs = self._model(ray_bundle)
        metrics_dict = self.model.get_metrics_dict(model_outputs, batch)
        loss_dict = self.model.get_loss_dict(model_outputs, batch, metrics_dict)
        return model_outputs, loss_dict, metrics_dict


@dataclass
class VanillaPipelineConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: VanillaPipeline)
    datamanager: Any = field(default_factory=VanillaDataManagerConfig)
    model: Any = field(default_factory=ModelConfig)

    def setup(self, **kwargs) -> Any:
        return self._target(self, **kwargs)


@dataclass
class TrainerConfig:
    method_name: str = ""
    steps_per_eval_batch: int = 500
    steps_per_save: int = 1000
    max_num_iterations: int = 1000000
    mixed_precision: bool = False
    pipeline: Any = None
    optimizers: Dict[str, Any] = field(default_factory=dict)
    viewer: Any = None
    vis: str = "wandb"


@dataclass
class MethodSpecification:
    config: TrainerConfig
    description: str


@dataclass
class Frustums:
    origins: Any
    directions: Any
    starts: Any
    ends: Any
    pixel_area: Any
    offsets: Any = None


@dataclass
class RaySamples:
    frustums: Frustums
    camera_indices: Any = None
    deltas: Any = None
    spacing_starts: Any = None
    spacing_ends: Any = None
    spacing_to_euclidean_fn: Optional[Callable] = None
    metadata: Any = None
    times: Any = None


@dataclass
class RayBundle:
    origins: Any
    directions: Any
    pixel_area: Any
    camera_indices: Any = None
    nears: Any = None
    fars: Any = None
    metadata: Any = None
    times: Any = None


def install() -> None:
    if importlib.util.find_spec("nerfstudio") is not None and "nerfstudio" not in sys.modules:
        raise RuntimeError("a real nerfstudio is installed; use it instead of the stand-in")
    _mod("nerfstudio", __rsn_standin__=True)
    _mod("nerfstudio.cameras")
    _mod("nerfstudio.cameras.rays", RayBundle=RayBundle, RaySamples=RaySamples, Frustums=Frustums)
    _mod("nerfstudio.configs")
    _mod("nerfstudio.configs.base_config", InstantiateConfig=InstantiateConfig, ViewerConfig=ViewerConfig)
    _mod("nerfstudio.data")
    _mod("nerfstudio.data.dataparsers")
    _mod("nerfstudio.data.dataparsers.blender_dataparser", BlenderDataParserConfig=BlenderDataParserConfig)
    _mod("nerfstudio.data.datamanagers")
    _mod("nerfstudio.data.datamanagers.base_datamanager", DataManagerConfig=DataManagerConfig,
         VanillaDataManager=VanillaDataManager, VanillaDataManagerConfig=VanillaDataManagerConfig)
    _mod("nerfstudio.engine")
    _mod("nerfstudio.engine.optimizers", OptimizerConfig=OptimizerConfig, RAdamOptimizerConfig=RAdamOptimizerConfig)
    _mod("nerfstudio.engine.schedulers", ExponentialDecaySchedulerConfig=ExponentialDecaySchedulerConfig)
    _mod("nerfstudio.engine.trainer", TrainerConfig=TrainerConfig)
    _mod("nerfstudio.models")
    _mod("nerfstudio.models.base_model", Model=Model, ModelConfig=ModelConfig)
    _mod("nerfstudio.pipelines")
    _mod("nerfstudio.pipelines.base_pipeline", VanillaPipeline=VanillaPipeline, VanillaPipelineConfig=VanillaPipelineConfig)
    _mod("nerfstudio.plugins")
    _mod("nerfstudio.plugins.types", MethodSpecification=MethodSpecification)
