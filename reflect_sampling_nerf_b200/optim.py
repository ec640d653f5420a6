"""Fused optimizer of the drop-in: torch.optim.Optimizer front end of rsn_radam_step.

Replaces what reflect_sampling_nerf_config.py:50-53 configures -- RAdamOptimizerConfig(lr=1e-3, eps=1e-15) +
ExponentialDecaySchedulerConfig(lr_final=1e-4, max_steps=50000) -- by, per step,
    rsn_radam_step (one launch over all 32 trained parameters: RAdam update, lr schedule from a device-side step counter)
    rsn_pack_field (one launch: the bf16 operand images the kernels stream follow the new fp32 master weights)
instead of torch's ~15 multi_tensor_apply launches, the LambdaLR host arithmetic and a lazy re-pack.  The gradients are
the flat vector train_path._flush_grads produced (p.grad are views of it).  Nothing here reads the device: the step is
CUDA-graph capturable.

Under nerfstudio, `FusedRAdamOptimizerConfig` (config.py) plugs it into the Trainer's Optimizers; the scheduler slot can
stay the reference's ExponentialDecay (then construct with lr_final=0 and the Trainer's LambdaLR drives `lr`), or be
dropped in favour of the in-kernel schedule.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
from torch import Tensor

from . import ops


class FusedRAdam(torch.optim.Optimizer):
    """RAdam over the parameters of ONE ReflectSamplingNeRFNerfField (`field`), fused."""

    def __init__(self, params: Iterable[Tensor], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-15,
                 weight_decay: float = 0.0, lr_final: float = 0.0, max_steps: int = 0, field=None) -> None:
        if weight_decay != 0.0:
            raise ValueError("FusedRAdam implements weight_decay = 0 (reflect_sampling_nerf_config.py:51)")
        if field is None:
            raise ValueError("FusedRAdam needs the field whose parameters it updates (field=model.field)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, lr_final=lr_final, max_steps=max_steps)
        super().__init__(params, defaults)
        self.field = field
        named = dict(field.named_parameters())
        mine = {id(p) for g in self.param_groups for p in g["params"]}
        missing = [k for k in ops.PACK_ORDER if id(named[k]) not in mine]
        if missing:
            raise ValueError(f"FusedRAdam: the parameter groups do not contain {missing[:3]}...")
        self._named = {k: named[k] for k in ops.PACK_ORDER}
        self._state_dev = None
        field.__dict__["_flat_grad_overwrite"] = True       # .grad stay views of the flat vector from step to step

    def _lazy_state(self, device):
        if self._state_dev is None or self._state_dev["exp_avg"].device != device:
            _, total = ops.flat_layout()
            self._state_dev = {
                "exp_avg": torch.zeros(total, device=device), "exp_avg_sq": torch.zeros(total, device=device),
                "step": torch.zeros(1, dtype=torch.int64, device=device),
                "counter": torch.zeros(1, dtype=torch.int32, device=device),
                "lr": torch.zeros(1, device=device),
            }
        return self._state_dev

    def zero_grad(self, set_to_none: bool = True) -> None:
        """No-op: every backward overwrites the flat gradient vector (p.grad are views of it) in full."""

    @torch.no_grad()
    def step(self, closure=None) -> Optional[Tensor]:
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        flat = self.field.__dict__.get("_flat_grad")
        if flat is None:
            raise RuntimeError("FusedRAdam.step(): no gradient -- run loss.backward() through the model first")
        st = self._lazy_state(flat.device)
        g = self.param_groups[0]
        ops.radam_step(self._named, flat, st["exp_avg"], st["exp_avg_sq"], st["step"], st["counter"], st["lr"],
                       lr=g["lr"], lr_final=g["lr_final"], max_steps=g["max_steps"], betas=g["betas"], eps=g["eps"])
        self.field.repack()
        return loss

    # -- checkpoints: the flat moment vectors + the device step counter -----------------------------------------------
    def state_dict(self):
        sd = super().state_dict()
        if self._state_dev is not None:
            sd["fused"] = {k: v.clone() for k, v in self._state_dev.items() if k != "counter"}
        return sd

    def load_state_dict(self, state_dict) -> None:
        fused = state_dict.pop("fused", None) if isinstance(state_dict, dict) else None
        super().load_state_dict(state_dict)
        if fused is not None:
            st = self._lazy_state(fused["exp_avg"].device)
            for k, v in fused.items():
                st[k].copy_(v)

    @property
    def steps_taken(self) -> Tensor:
        """Device int64 [1]."""
        return self._lazy_state(next(iter(self._named.values())).device)["step"]
