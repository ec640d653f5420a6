"""Drop-in field: same parameters, module paths, state-dict keys AND method contract as ReflectSamplingNeRFNerfField
(reflect_sampling_nerf_field.py:28-207), evaluated by the fused sm_100a kernels instead of eager PyTorch.

Two ways in:

* The model's passes (`evaluate_samples`, `get_inf_color`; train_path.py) -- ONE kernel launch per pass over all samples
  of a ray batch: get_blob -> contract -> get_density -> heads -> IDE -> get_mid fused (csrc/field_fwd.cu).
* The reference's method-by-method API (field.py:90-207): `get_blob`, `contract`, `get_density`, `get_pred_normals`,
  `get_normals`, `get_roughness`, `get_low`, `get_mid`, `get_diff`, `get_tint`, `get_inf_color`, `get_reflection`, with
  the reference's signatures.  `get_density` runs the fused kernel once for everything that hangs off the embedding and
  returns the density plus a `FieldEmbedding` handle; the head methods read their slice of that one evaluation.  (The
  256-wide embedding itself never leaves the SM: a handle stands in for the tensor the reference passes around.)
  When the Gaussians came from `get_blob(ray_samples)` the pass runs straight from the samples' bins -- the same launch
  as the model's -- otherwise on the caller's mean / cov tensors (rsn_field_forward_points).  No autograd on this route:
  training goes through the model (train_path.py), whose backward kernels need the fused pass structure.

Parameters stay fp32 nn.Parameters owned by PyTorch (optimizers, GradScaler, DDP and checkpoints keep working,
SURVEY.md §5); the bf16 operand blob the kernels stream is derived state, re-packed whenever a parameter changes.
"""
from __future__ import annotations

from typing import Optional, Tuple

import weakref

import torch
from torch import Tensor, nn

from . import ops


class _Layers(nn.Module):
    """nerfstudio MLP parameter container: `layers` = ModuleList of nn.Linear (keys `*.layers.{i}.weight`)."""

    def __init__(self, dims) -> None:
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(i, o) for i, o in dims])
        self._out_dim = dims[-1][1]

    def get_out_dim(self) -> int:
        return self._out_dim


class _Head(nn.Module):
    """nerfstudio FieldHead parameter container: `net` = nn.Linear (keys `*.net.weight`)."""

    def __init__(self, in_dim: int, out_dim: int) -> None:
        super().__init__()
        self.net = nn.Linear(in_dim, out_dim)


class FieldEmbedding:
    """Stands in for the [...,256] `mlp_out` tensor of field.py:132: everything the heads derive from it, evaluated once."""

    def __init__(self, lead_shape, sigma, feat, aux, mean, cov, dirs, normals=None) -> None:
        self.lead_shape = tuple(lead_shape)
        self.sigma, self.feat, self.aux = sigma, feat, aux            # [P], [P,16], [P,8]
        self.mean, self.cov, self.dirs = mean, cov, dirs              # what the evaluation consumed ([P,*]) -- for re-runs
        self.normals = normals

    @property
    def shape(self):
        return torch.Size(self.lead_shape + (256,))

    @property
    def device(self):
        return self.feat.device

    def _col(self, sl) -> Tensor:
        out = self.feat[:, sl]
        return out.reshape(*self.lead_shape, -1)


class ReflectSamplingNeRFNerfField(nn.Module):
    """B200-native field.  Constructor arguments mirror field.py:36-48; the architecture the kernels implement
    is the one the reference model instantiates (model.py:98-106): 99-dim IPE, 8x256 base MLP with a skip at
    layer 4, 34-dim IDE, one 128-wide mid layer."""

    ENC_DIM, IDE_DIM, WIDTH, MID_WIDTH = 99, 34, 256, 128

    def __init__(self, position_encoding=None, direction_encoding=None, base_mlp_num_layers: int = 8,
                 base_mlp_layer_width: int = 256, skip_connections: Tuple[int, ...] = (4,),
                 head_mlp_num_layers: int = 1, head_mlp_layer_width: int = 128, spatial_distortion=None,
                 density_bias: float = 0.5, roughness_bias: float = -1.0) -> None:
        super().__init__()
        if (base_mlp_num_layers, base_mlp_layer_width, tuple(skip_connections), head_mlp_num_layers,
                head_mlp_layer_width) != (8, 256, (4,), 1, 128):
            raise ValueError("the fused sm_100a field kernels implement the 8x256 (skip 4) + 1x128 architecture "
                             "that reflect_sampling_nerf_model.py:98-106 instantiates")
        if spatial_distortion is not None:
            raise ValueError("spatial_distortion is always None in the reference (model.py:103-106)")
        if density_bias != 0.5:
            raise ValueError("density_bias is compiled into the kernels as 0.5 (field.py:46)")
        for enc, dim, what in ((position_encoding, self.ENC_DIM, "position"), (direction_encoding, self.IDE_DIM, "direction")):
            if enc is not None and enc.get_out_dim() != dim:
                raise ValueError(f"the fused kernels implement the {dim}-dim {what} encoding of model.py:98-101")
        from .components import IntegratedSHEncoding, NeRFEncoding
        # (field.py:49-51; the fused kernels evaluate these encodings in their prologue / epilogue, the modules serve the
        # component API and direct tests)
        self.position_encoding = position_encoding if position_encoding is not None else NeRFEncoding()
        self.direction_encoding = direction_encoding if direction_encoding is not None else IntegratedSHEncoding()
        self.spatial_distortion = spatial_distortion
        w, e = self.WIDTH, self.ENC_DIM
        # construction order = field.py:54-86, so a seeded init reproduces the reference's parameters
        self.mlp_base = _Layers([(e, w)] + [(w + e if i == 4 else w, w) for i in range(1, 8)])
        self.field_output_density = _Head(w, 1)
        self.density_bias = density_bias
        self.softplus = nn.Softplus()
        self.sigmoid = nn.Sigmoid()
        self.field_output_low = _Head(w, 3)          # present, never used
        self.field_output_bottleneck = _Head(w, w)
        self.mlp_mid = _Layers([(self.IDE_DIM + w, self.MID_WIDTH)])
        self.field_output_mid = _Head(self.MID_WIDTH, 3)
        self.field_output_normals = _Head(w, 3)
        self.field_output_roughness = _Head(w, 1)
        self.roughness_bias = roughness_bias         # stored, never used (App. B Q2)
        self.field_output_diff = _Head(w, 3)
        self.field_output_tint = _Head(w, 3)
        self._packed: Optional[Tuple[Tensor, Tensor]] = None
        self._packed_key = None
        self._packed_t: Optional[Tuple[Tensor, Tensor]] = None
        # training state (see train_path.py): gradient blob of the wgrad kernel for the backward in flight,
        # the reusable dY stash, and the data-parallel world size for the flat gradient all-reduce
        self._grad_blob: Optional[Tensor] = None
        self._dy_buffer: Optional[Tensor] = None
        self.dp_world_size = 1
        self._normals: Optional[Tensor] = None       # density-gradient normals of the last get_density(..., True)
        ReflectSamplingNeRFNerfField._instances.add(self)

    _instances: "weakref.WeakSet" = None  # type: ignore[assignment]

    @classmethod
    def owner_of(cls, params) -> "ReflectSamplingNeRFNerfField":
        """The live field whose parameters `params` are (an optimizer config only receives the parameter list)."""
        ids = {id(p) for p in params}
        for f in list(cls._instances):
            if all(id(p) in ids for p in f.parameters()):
                return f
        raise ValueError("no ReflectSamplingNeRFNerfField owns these parameters")

    # ------------------------------------------------------------------------------------ derived state
    def _version_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _repack(self) -> None:
        key = self._version_key()
        if self._packed is None or key != self._packed_key:
            self.repack()

    def repack(self) -> None:
        """bf16 operand images <- current fp32 parameters (one kernel, csrc/pack.cu).  Existing blobs are overwritten in
        place, so their device addresses are stable (CUDA graphs; the fused optimizer calls this after every step)."""
        out = None
        dev = self.mlp_base.layers[0].weight.device
        if self._packed is not None and self._packed[0].device == dev:
            out = (self._packed[0], self._packed[1], self._packed_t[0], self._packed_t[1])
        with torch.no_grad():
            wblob, bias, wblob_t, wd = ops.pack_field(dict(self.named_parameters()), out)
        self._packed, self._packed_t, self._packed_key = (wblob, bias), (wblob_t, wd), self._version_key()

    def packed(self) -> Tuple[Tensor, Tensor]:
        """(bf16 operand blob, fp32 bias vector) for the current parameter values."""
        self._repack()
        return self._packed

    def packed_t(self) -> Tuple[Tensor, Tensor]:
        """(transposed bf16 operand blob of the dgrad chains, bf16 density-head row)."""
        self._repack()
        return self._packed_t

    # ------------------------------------------------------------------------------------ the model's passes
    def evaluate_samples(self, origins: Tensor, directions: Tensor, pixel_area: Tensor, bins: Tensor,
                         count: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """One fused pass over every frustum sample of a ray batch (field.py:90-186 + components.py:52-140).
        origins/directions [N,3], pixel_area [N,1], bins [N,S+1] -> density [N,S], feat [N,S,16]."""
        wblob, bias = self.packed()
        return ops.field_forward(wblob, bias, origins, directions, pixel_area, bins, count)

    # ------------------------------------------------------------------------------------ the reference's method API
    def get_blob(self, ray_samples) -> Tuple[Tensor, Tensor]:
        """field.py:90-96."""
        bins = getattr(ray_samples, "_rsn_bins", None)
        if bins is not None and getattr(ray_samples, "_rsn_bundle", None) is not None:
            b = ray_samples._rsn_bundle
            mean, cov = ops.frustum_gaussians(b.origins, b.directions, b.pixel_area, bins[1])
        else:
            g = ray_samples.frustums.get_gaussian_blob()
            mean, cov = g.mean, g.cov
        if self.spatial_distortion is not None:  # pragma: no cover - always None (constructor)
            raise ValueError("spatial_distortion is not supported")
        mean._rsn_link = (ray_samples, False)
        return mean, cov

    def contract(self, mean: Tensor, cov: Tensor, mask_return: bool = False):
        """field.py:98-119."""
        mean_c, cov_c = ops.contract(mean, cov)
        link = getattr(mean, "_rsn_link", None)
        if link is not None:
            mean_c._rsn_link = (link[0], True)
            mean_c._rsn_cov = cov_c
        if mask_return:
            return mean_c, cov_c, torch.linalg.norm(mean, dim=-1, keepdim=True) > 1
        return mean_c, cov_c

    @torch.no_grad()
    def get_density(self, mean: Tensor, cov: Optional[Tensor] = None, requires_density_grad: bool = False):
        """field.py:122-137 -> (density [...,1], FieldEmbedding)."""
        wblob, bias = self.packed()
        lead = mean.shape[:-1]
        link = getattr(mean, "_rsn_link", None)
        linked = (link is not None and link[1] and cov is not None and getattr(mean, "_rsn_cov", None) is cov
                  and getattr(link[0], "_rsn_bins", None) is not None and getattr(link[0], "_rsn_bundle", None) is not None)
        normals = None
        if linked:
            rs = link[0]
            b, eu = rs._rsn_bundle, rs._rsn_bins[1]
            if requires_density_grad and self.training:      # field.py:125-127,134-135 -> get_normals()
                from .train_path import _claim_stash
                slot = _claim_stash(self, eu.shape[0] * (eu.shape[1] - 1), eu.device)
                try:
                    sigma, feat, stash, aux = ops.field_forward_train(wblob, bias, ops.MODE_SAMPLES, b.origins, b.directions,
                                                                      b.pixel_area, eu, slot.buf)
                    wblob_t, wd = self.packed_t()
                    normals = ops.field_normals(wblob_t, wd, stash, sigma.shape[0], sigma.shape[1])
                finally:
                    slot.in_flight = False
            else:
                sigma, feat, aux = ops.field_forward(wblob, bias, b.origins, b.directions, b.pixel_area, eu, want_aux=True)
            dirs = b.directions[:, None, :].expand(*sigma.shape, 3)
            sigma, feat, aux = sigma.reshape(-1), feat.reshape(-1, 16), aux.reshape(-1, 8)
            dirs_flat = dirs.reshape(-1, 3)
        else:
            if requires_density_grad and self.training:
                raise NotImplementedError(
                    "density-gradient normals need the samples the Gaussians came from: pass the tensors returned by "
                    "get_blob(ray_samples) -> contract(...) unchanged (reflect_sampling_nerf_model.py:151-153)")
            m = mean.reshape(-1, 3)
            c = cov.reshape(-1, 3, 3) if cov is not None else m.new_zeros(m.shape[0], 3, 3)
            dirs_flat = m.new_zeros(m.shape[0], 3)          # the view direction arrives with get_mid
            sigma, feat, aux = ops.field_forward_points(wblob, bias, m, c, dirs_flat)
        self._normals = None if normals is None else normals.reshape(*lead, 3)
        emb = FieldEmbedding(lead, sigma, feat, aux, mean.reshape(-1, 3), None if cov is None else cov.reshape(-1, 3, 3),
                             dirs_flat, self._normals)
        emb.has_dirs = linked
        return sigma.reshape(*lead, 1), emb

    @staticmethod
    def _emb(embedding) -> FieldEmbedding:
        if not isinstance(embedding, FieldEmbedding):
            raise TypeError("expected the embedding handle returned by get_density (the 256-wide embedding tensor is "
                            "never materialised by the fused kernels)")
        return embedding

    def get_pred_normals(self, embedding) -> Tensor:
        """field.py:139-144.  (Evaluated with the view direction only for n.d; the normal itself does not depend on it.)"""
        return self._emb(embedding)._col(ops.F_NORMAL)

    def get_normals(self) -> Tensor:
        """field.py:146-147: -normalize(d raw density / d contracted mean) of the last get_density(..., True)."""
        if self._normals is None:
            raise RuntimeError("get_normals() needs a preceding get_density(mean, cov, requires_density_grad=True) in "
                               "training mode (field.py:125-127)")
        return self._normals

    def get_roughness(self, embedding, activation: Optional[nn.Module] = None) -> Tensor:
        """field.py:150-155 (default activation: Sigmoid)."""
        e = self._emb(embedding)
        if activation is None or isinstance(activation, nn.Sigmoid):
            return e._col(slice(ops.F_ROUGH_SIGMOID, ops.F_ROUGH_SIGMOID + 1))
        if isinstance(activation, nn.Softplus) and activation.beta == 1 and activation.threshold == 20:
            return e._col(slice(ops.F_ROUGH_SOFTPLUS, ops.F_ROUGH_SOFTPLUS + 1))
        return activation(e.aux[:, 6:7].reshape(*e.lead_shape, 1))

    def _rerun(self, e: FieldEmbedding, dirs: Tensor, rho: Optional[Tensor]):
        if e.cov is None:
            cov = e.mean.new_zeros(e.mean.shape[0], 3, 3)
        else:
            cov = e.cov
        wblob, bias = self.packed()
        return ops.field_forward_points(wblob, bias, e.mean, cov, dirs, rho)

    @torch.no_grad()
    def get_mid(self, directions: Tensor, roughness: Tensor, embedding, use_bottleneck: bool = True) -> Tensor:
        """field.py:167-174: sigmoid(Linear(ReLU(Linear([IDE(directions, roughness), bottleneck(embedding)]))))."""
        if not use_bottleneck:
            raise NotImplementedError("use_bottleneck=False is never used by the reference (model.py:174,208,309,335)")
        e = self._emb(embedding)
        d = directions.reshape(-1, 3)
        rho = roughness.reshape(-1)
        same_dirs = getattr(e, "has_dirs", False) and d.shape == e.dirs.shape and (
            d.data_ptr() == e.dirs.data_ptr() or bool(torch.equal(d, e.dirs)))
        sp = e.feat[:, ops.F_ROUGH_SOFTPLUS]
        same_rho = rho.shape == sp.shape and (rho.data_ptr() == sp.data_ptr() or bool(torch.equal(rho, sp)))
        if same_dirs and same_rho:
            return e.aux[:, 0:3].reshape(*e.lead_shape, 3)
        _, _, aux = self._rerun(e, d, rho)
        return aux[:, 0:3].reshape(*e.lead_shape, 3)

    @torch.no_grad()
    def get_low(self, embedding, use_bottleneck: bool = True) -> Tensor:
        """field.py:158-164: the mid colour with a zero direction encoding (roughness -> infinity damps every IDE band)."""
        if not use_bottleneck:
            raise NotImplementedError("use_bottleneck=False is never used by the reference")
        e = self._emb(embedding)
        _, _, aux = self._rerun(e, e.dirs, torch.full_like(e.sigma, 1e30))
        return aux[:, 0:3].reshape(*e.lead_shape, 3)

    def get_diff(self, embedding) -> Tensor:
        """field.py:176-180."""
        return self._emb(embedding)._col(ops.F_DIFF)

    def get_tint(self, embedding) -> Tensor:
        """field.py:182-186."""
        return self._emb(embedding)._col(ops.F_TINT)

    def get_inf_color(self, directions: Tensor, sqradius: Tensor, count: Optional[Tensor] = None) -> Tensor:
        """field.py:190-201."""
        wblob, bias = self.packed()
        lead = directions.shape[:-1]
        out = ops.field_inf_color(wblob, bias, directions.reshape(-1, 3), sqradius.reshape(-1), count)
        return out.reshape(*lead, 3)

    def get_reflection(self, directions: Tensor, normals: Tensor) -> Tuple[Tensor, Tensor]:
        """field.py:203-207 (per-sample elementwise; inside the fused pass n.d is feature column 13)."""
        n_dot_d = torch.sum(directions * normals, dim=-1, keepdim=True)
        reflections = directions - 2 * n_dot_d * normals
        reflections = torch.nn.functional.normalize(reflections, dim=-1)
        return reflections, n_dot_d


ReflectSamplingNeRFNerfField._instances = weakref.WeakSet()
