"""Python-side operators over the C-ABI (one function / autograd.Function per kernel stage).

Tensors stay torch-owned; these wrappers only check shapes, allocate outputs and pass raw device
pointers + the current CUDA stream to `librsn_b200.so`.  No CPU path exists.
"""
from __future__ import annotations

import functools
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib

UNIFORM, RECIPROCAL = 0, 1
# bench.py sets PROFILE to a list to collect (kernel, start event, end event, algorithmic FLOPs) per launch of
# the dominant kernel on the launching stream; None (default) = no events recorded.
PROFILE = None
FLOP_PER_POINT = {0: 1230592, 1: 1225472}   # SURVEY.md §8d (forward; primary figure used for every sample pass)
# algorithmic (FLOPs, HBM bytes) per point of the four field kernels
KERNEL_WORK = {
    "field_fwd_kernel": (1230592, 68),
    "field_fwd_kernel[train]": (1230592, 68 + 32 + 37 * 128 + 288),      # 41 stash blocks, the 4 bottleneck ones not written
    "field_chain_kernel<normals>": (1019392, 288 + 4 * 128 + 12),
    "field_chain_kernel<backward>": (1179904, 288 + 35 * 128 + 160),         # 39 dY blocks, the 4 of dY_bott not written
    "field_chain_kernel<backward+area>": (1229056, 288 + 4 * 128 + 35 * 128 + 164),
    # every stash block that exists once: 37 activation + 35 dY blocks of 128 B per point (the 12 active jobs issue 82 block
    # reads per tile: 10 operands are read by two jobs)
    "field_wgrad_kernel": (1230592, 72 * 128),
    # fused backward: dgrad + wgrad FLOPs; HBM: masks + dY written once + X read (dY read back from L2)
    "field_bwd_fused_kernel": (1179904 + 1230592, 288 + 35 * 128 + 160 + 52 * 128),
    "field_bwd_fused_kernel+area": (1229056 + 1230592, 288 + 4 * 128 + 35 * 128 + 164 + 52 * 128),
    # K8 at C = 16 channels, per SAMPLE: sigma 4 + bin 4 + feat 64 in, weight 4 out | + dL/dw 4 in, dL/dsigma 4 + dL/dfeat 64 out
    "composite_fwd_kernel": (0, 76),
    "composite_bwd_kernel": (0, 144),
}


class _Prof:
    """CUDA-event bracket around one kernel launch on the launching stream, active only while PROFILE is a list."""

    def __init__(self, name: str, n_points: int):
        self.name, self.n, self.on = name, n_points, PROFILE is not None

    def __enter__(self):
        if self.on:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.on and PROFILE is not None:
            self.e1.record()
            flop, nbytes = KERNEL_WORK[self.name]
            PROFILE.append((self.name, self.e0, self.e1, self.n * flop, self.n * nbytes))
        return False


def _f32c(t: Optional[Tensor]) -> Optional[Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


@functools.lru_cache(maxsize=32)
def _linspace_bins(n_samples: int, device: torch.device) -> Tensor:
    # table taken from torch itself so it is bit-identical to SpacedSampler's
    return torch.linspace(0.0, 1.0, n_samples + 1).to(device)


@functools.lru_cache(maxsize=32)
def _pdf_u_base(n_out: int, train: bool, device: torch.device) -> Tensor:
    nb = n_out + 1
    u = torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb)
    if not train:
        u = u + 1.0 / (2 * nb)
    return u.to(device)


# ----------------------------------------------------------------------------------------- K1
def sample_spaced(nears: Tensor, fars: Tensor, n_samples: int, kind: int,
                  t_rand: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """-> (spacing bins [N,S+1], euclidean bins [N,S+1]).  t_rand: [N,S+1] or [N,1] in [0,1), None = eval."""
    nears, fars, t_rand = _f32c(nears.reshape(-1)), _f32c(fars.reshape(-1)), _f32c(t_rand)
    n = nears.shape[0]
    dev = nears.device
    spacing = torch.empty(n, n_samples + 1, device=dev, dtype=torch.float32)
    euclid = torch.empty_like(spacing)
    cols = 0 if t_rand is None else t_rand.shape[-1]
    if t_rand is not None and t_rand.shape[0] != n:
        raise ValueError(f"t_rand has {t_rand.shape[0]} rows for {n} rays")
    _lib.call("rsn_sample_spaced", _lib.ptr(nears), _lib.ptr(fars), _lib.ptr(_linspace_bins(n_samples, dev)),
              _lib.ptr(t_rand), cols, kind, _lib.ptr(spacing), _lib.ptr(euclid), n, n_samples, _lib.stream())
    return spacing, euclid


# ----------------------------------------------------------------------------------------- K2
def pdf_resample(weights: Tensor, spacing_bins: Tensor, nears: Tensor, fars: Tensor, n_out: int, kind: int,
                 rand: Optional[Tensor] = None, train: Optional[bool] = None, histogram_padding: float = 0.01,
                 return_inds: bool = False):
    """weights [N,S] (or [N,S,1]); spacing_bins [N,S+1] -> (spacing [N,n_out+1], euclid [N,n_out+1][, inds])."""
    if weights.dim() == 3:
        weights = weights[..., 0]
    weights, spacing_bins = _f32c(weights.detach()), _f32c(spacing_bins)
    nears, fars, rand = _f32c(nears.reshape(-1)), _f32c(fars.reshape(-1)), _f32c(rand)
    n, s = weights.shape
    if spacing_bins.shape != (n, s + 1):
        raise ValueError(f"spacing_bins must be [{n},{s + 1}], got {tuple(spacing_bins.shape)}")
    if train is None:
        train = rand is not None
    if train and rand is None:
        rand = torch.rand(n, n_out + 1, device=weights.device)
    if rand is not None and rand.shape != (n, n_out + 1):
        raise ValueError(f"rand must be [{n},{n_out + 1}], got {tuple(rand.shape)}")
    dev = weights.device
    out_s = torch.empty(n, n_out + 1, device=dev, dtype=torch.float32)
    out_e = torch.empty_like(out_s)
    inds = torch.empty(n, n_out + 1, device=dev, dtype=torch.int64) if return_inds else None
    _lib.call("rsn_pdf_resample", _lib.ptr(weights), weights.stride(0), _lib.ptr(spacing_bins), _lib.ptr(nears),
              _lib.ptr(fars), _lib.ptr(_pdf_u_base(n_out, bool(train), dev)), _lib.ptr(rand), kind,
              float(histogram_padding), _lib.ptr(out_s), _lib.ptr(out_e), _lib.ptr(inds), n, s, n_out,
              _lib.stream())
    return (out_s, out_e, inds) if return_inds else (out_s, out_e)


# ----------------------------------------------------------------------------------------- K8
_COMPOSITE_CHANNELS = (0, 1, 3, 4, 8, 16)


class _Composite(torch.autograd.Function):
    """sigma [N,S], bins [N,S+1] (euclidean, no grad), feat [N,S,C] or None
    -> weights [N,S], accumulation [N], median depth [N] (no grad), feat_out [N,C]."""

    @staticmethod
    def forward(ctx, sigma: Tensor, bins: Tensor, feat: Optional[Tensor]):
        sigma, bins, feat = _f32c(sigma), _f32c(bins), _f32c(feat)
        n, s = sigma.shape
        c = 0 if feat is None else feat.shape[-1]
        if c not in _COMPOSITE_CHANNELS:
            raise ValueError(f"composite: channel count {c} not in {_COMPOSITE_CHANNELS}")
        if bins.shape != (n, s + 1):
            raise ValueError(f"composite: bins must be [{n},{s + 1}], got {tuple(bins.shape)}")
        dev = sigma.device
        weights = torch.empty(n, s, device=dev, dtype=torch.float32)
        acc = torch.empty(n, device=dev, dtype=torch.float32)
        depth = torch.empty(n, device=dev, dtype=torch.float32)
        feat_out = torch.empty(n, c, device=dev, dtype=torch.float32)
        with _Prof("composite_fwd_kernel", n * s if c == 16 else 0):
            _lib.call("rsn_composite_fwd", _lib.ptr(sigma), bins.data_ptr(), bins.data_ptr() + 4, s + 1,
                      _lib.ptr(feat), c, _lib.ptr(weights), _lib.ptr(acc), _lib.ptr(depth),
                      _lib.ptr(feat_out) if c else None, n, s, _lib.stream())
        ctx.save_for_backward(sigma, bins, feat if feat is not None else sigma.new_empty(0))
        ctx.c = c
        ctx.mark_non_differentiable(depth)
        return weights, acc, depth, feat_out

    @staticmethod
    def backward(ctx, g_w, g_acc, _g_depth, g_feat_out):
        sigma, bins, feat = ctx.saved_tensors
        c = ctx.c
        n, s = sigma.shape
        need_feat = c > 0 and ctx.needs_input_grad[2]
        g_sigma = torch.empty_like(sigma)
        g_feat = torch.empty_like(feat) if need_feat else None
        g_w, g_acc, g_feat_out = _f32c(g_w), _f32c(g_acc), _f32c(g_feat_out) if c else None
        with _Prof("composite_bwd_kernel", n * s if c == 16 else 0):
            _lib.call("rsn_composite_bwd", _lib.ptr(sigma), bins.data_ptr(), bins.data_ptr() + 4, s + 1,
                      _lib.ptr(feat) if c else None, c, _lib.ptr(g_w), _lib.ptr(g_acc), _lib.ptr(g_feat_out),
                      _lib.ptr(g_sigma), _lib.ptr(g_feat), n, s, _lib.stream())
        return g_sigma, None, g_feat


class _Composite16(torch.autograd.Function):
    """The model's form: 16 feature channels + the two per-sample normal losses fused in (rsn_composite16_*).
    sigma [N,S], bins [N,S+1], feat [N,S,16], normals [N,S,3] (constant)
    -> weights [N,S], accumulation [N], median depth [N], feat_out [N,16], pred_normal_loss [N], orientation_loss [N]"""

    @staticmethod
    def forward(ctx, sigma: Tensor, bins: Tensor, feat: Tensor, normals: Tensor):
        sigma, bins, feat, normals = _f32c(sigma), _f32c(bins), _f32c(feat), _f32c(normals)
        n, s = sigma.shape
        if feat.shape != (n, s, 16) or normals.shape != (n, s, 3) or bins.shape != (n, s + 1):
            raise ValueError("composite16: expects sigma [N,S], bins [N,S+1], feat [N,S,16], normals [N,S,3]")
        dev = sigma.device
        weights = torch.empty(n, s, device=dev, dtype=torch.float32)
        acc, depth, pnl, ol = (torch.empty(n, device=dev, dtype=torch.float32) for _ in range(4))
        feat_out = torch.empty(n, 16, device=dev, dtype=torch.float32)
        with _Prof("composite_fwd_kernel", n * s):
            _lib.call("rsn_composite16_fwd", _lib.ptr(sigma), bins.data_ptr(), bins.data_ptr() + 4, s + 1, _lib.ptr(feat),
                      _lib.ptr(normals), _lib.ptr(weights), _lib.ptr(acc), _lib.ptr(depth), _lib.ptr(feat_out),
                      _lib.ptr(pnl), _lib.ptr(ol), n, s, _lib.stream())
        ctx.save_for_backward(sigma, bins, feat, normals)
        ctx.mark_non_differentiable(depth)
        return weights, acc, depth, feat_out, pnl, ol

    @staticmethod
    def backward(ctx, g_w, g_acc, _g_depth, g_feat_out, g_pnl, g_ol):
        sigma, bins, feat, normals = ctx.saved_tensors
        n, s = sigma.shape
        g_sigma, g_feat = torch.empty_like(sigma), torch.empty_like(feat)
        # keep the contiguous copies alive until the launch is enqueued (a freed temporary's block is handed to the
        # next allocation at once)
        g_w, g_acc, g_feat_out, g_pnl, g_ol = (_f32c(t) for t in (g_w, g_acc, g_feat_out, g_pnl, g_ol))
        with _Prof("composite_bwd_kernel", n * s):
            _lib.call("rsn_composite16_bwd", _lib.ptr(sigma), bins.data_ptr(), bins.data_ptr() + 4, s + 1, _lib.ptr(feat),
                      _lib.ptr(normals), _lib.ptr(g_w), _lib.ptr(g_acc), _lib.ptr(g_feat_out), _lib.ptr(g_pnl),
                      _lib.ptr(g_ol), _lib.ptr(g_sigma), _lib.ptr(g_feat), n, s, _lib.stream())
        return g_sigma, None, g_feat, None


def composite16(sigma: Tensor, bins: Tensor, feat: Tensor, normals: Tensor):
    return _Composite16.apply(sigma, bins, feat, normals)


def composite(sigma: Tensor, bins: Tensor, feat: Optional[Tensor] = None):
    """Alpha compositing of one ray batch (K8).  Returns (weights, accumulation, median_depth, feat_out)."""
    return _Composite.apply(sigma, bins, feat)


# ----------------------------------------------------------------------------------------- K3+K4+K5+K7
MODE_SAMPLES, MODE_INF_COLOR = 0, 1
N_FEAT = 16
(F_RGB, F_DIFF, F_TINT, F_NORMAL, F_ROUGH_SIGMOID, F_NDOTD, F_RAW_DENSITY, F_ROUGH_SOFTPLUS) = (
    slice(0, 3), slice(3, 6), slice(6, 9), slice(9, 12), 12, 13, 14, 15)


def field_forward(wblob: Tensor, bias: Tensor, origins: Tensor, dirs: Tensor, pixel_area: Tensor,
                  bins: Tensor) -> Tuple[Tensor, Tensor]:
    """Fused field evaluation of every frustum sample of a ray batch (inference form, no autograd).
    origins/dirs [N,3], pixel_area [N] or [N,1], bins [N,S+1] euclidean -> sigma [N,S], feat [N,S,16]."""
    origins, dirs, bins = _f32c(origins), _f32c(dirs), _f32c(bins)
    area = _f32c(pixel_area.reshape(-1))
    n, s = bins.shape[0], bins.shape[1] - 1
    if origins.shape != (n, 3) or dirs.shape != (n, 3) or area.shape[0] != n:
        raise ValueError("field_forward: origins/dirs must be [N,3] and pixel_area [N] for bins [N,S+1]")
    sigma = torch.empty(n, s, device=bins.device, dtype=torch.float32)
    feat = torch.empty(n, s, N_FEAT, device=bins.device, dtype=torch.float32)
    with _Prof("field_fwd_kernel", n * s):
        _lib.call("rsn_field_forward", _lib.ptr(wblob), _lib.ptr(bias), MODE_SAMPLES, _lib.ptr(origins),
                  _lib.ptr(dirs), _lib.ptr(area), _lib.ptr(bins), n, s, _lib.ptr(sigma), _lib.ptr(feat), _lib.stream())
    return sigma, feat


def field_inf_color(wblob: Tensor, bias: Tensor, dirs: Tensor, sqradius: Tensor) -> Tensor:
    """field.get_inf_color (field.py:190-201): dirs [M,3], sqradius [M] or [M,1] -> rgb [M,3]."""
    dirs, sq = _f32c(dirs), _f32c(sqradius.reshape(-1))
    m = dirs.shape[0]
    sigma = torch.empty(m, device=dirs.device, dtype=torch.float32)
    feat = torch.empty(m, N_FEAT, device=dirs.device, dtype=torch.float32)
    _lib.call("rsn_field_forward", _lib.ptr(wblob), _lib.ptr(bias), MODE_INF_COLOR, None, _lib.ptr(dirs),
              _lib.ptr(sq), None, m, 1, _lib.ptr(sigma), _lib.ptr(feat), _lib.stream())
    return feat[:, :3]


# ----------------------------------------------------------------------------------------- training kernels
import ctypes as _ct  # noqa: E402


@functools.lru_cache(maxsize=1)
def wgrad_layout():
    """(offsets [2*J], shapes [(rows, cols)]*J, total floats) of the rsn_field_wgrad gradient blob."""
    offs = (_ct.c_int64 * 64)()
    shp = (_ct.c_int64 * 64)()
    tot = _ct.c_int64(0)
    n = _lib.lib().rsn_field_wgrad_layout(offs, shp, _ct.byref(tot))
    return [int(offs[i]) for i in range(2 * n)], [(int(shp[2 * j]), int(shp[2 * j + 1])) for j in range(n)], int(tot.value)


def field_forward_train(wblob: Tensor, bias: Tensor, mode: int, origins: Optional[Tensor], dirs: Tensor,
                        area: Tensor, bins: Optional[Tensor], stash: Optional[Tensor] = None):
    """Forward pass that also leaves the activation stash + aux for the backward kernels.
    mode 0: bins [N,S+1]; mode 1 (infinity colour): one point per ray.  -> sigma [N,S], feat [N,S,16], stash, aux"""
    dirs, area = _f32c(dirs), _f32c(area.reshape(-1))
    n = dirs.shape[0]
    if mode == MODE_SAMPLES:
        origins, bins = _f32c(origins), _f32c(bins)
        s = bins.shape[1] - 1
    else:
        s = 1
    dev = dirs.device
    sigma = torch.empty(n, s, device=dev, dtype=torch.float32)
    feat = torch.empty(n, s, N_FEAT, device=dev, dtype=torch.float32)
    aux = torch.empty(n, s, 8, device=dev, dtype=torch.float32)
    nbytes = _lib.lib().rsn_field_stash_bytes(n * s)
    if stash is None:
        stash = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    elif stash.numel() < nbytes or stash.dtype != torch.uint8 or stash.device != dev:
        raise ValueError("field_forward_train: stash workspace too small")
    with _Prof("field_fwd_kernel[train]", n * s):
        _lib.call("rsn_field_forward_train", _lib.ptr(wblob), _lib.ptr(bias), mode, _lib.ptr(origins), _lib.ptr(dirs),
                  _lib.ptr(area), _lib.ptr(bins), n, s, _lib.ptr(sigma), _lib.ptr(feat), _lib.ptr(stash),
                  _lib.ptr(aux), _lib.stream())
    return sigma, feat, stash, aux


def field_normals(wblob_t: Tensor, wd_bf16: Tensor, stash: Tensor, n: int, s: int) -> Tensor:
    """K6: -normalize(d raw_density / d contracted mean) of every sample of the pass that produced `stash`."""
    normals = torch.empty(n, s, 3, device=stash.device, dtype=torch.float32)
    with _Prof("field_chain_kernel<normals>", n * s):
        _lib.call("rsn_field_normals", _lib.ptr(wblob_t), _lib.ptr(wd_bf16), _lib.ptr(stash), n, s,
                  _lib.ptr(normals), _lib.stream())
    return normals


def field_backward(wblob_t: Tensor, stash: Tensor, mode: int, origins, dirs, area, bins, n: int, s: int,
                   g_sigma: Optional[Tensor], g_feat: Tensor, feat: Tensor, aux: Tensor, dy_stash: Tensor,
                   want_area: bool) -> Optional[Tensor]:
    """K5 dgrad chain: fills dy_stash; returns dL/d pixel_area (mode 0) / dL/d sqradius (mode 1) per POINT [n,s]
    when want_area."""
    g_area = torch.empty(n, s, device=stash.device, dtype=torch.float32) if want_area else None
    with _Prof("field_chain_kernel<backward+area>" if want_area else "field_chain_kernel<backward>", n * s):
        _lib.call("rsn_field_backward", _lib.ptr(wblob_t), _lib.ptr(stash), mode, _lib.ptr(origins), _lib.ptr(dirs),
                  _lib.ptr(area), _lib.ptr(bins), n, s, _lib.ptr(g_sigma), _lib.ptr(g_feat), _lib.ptr(feat),
                  _lib.ptr(aux), _lib.ptr(dy_stash), _lib.ptr(g_area), _lib.stream())
    return g_area


def field_backward_fused(wblob_t: Tensor, stash: Tensor, mode: int, origins, dirs, area, bins, n: int, s: int,
                         g_sigma: Optional[Tensor], g_feat: Tensor, feat: Tensor, aux: Tensor, dy_stash: Tensor,
                         want_area: bool, grad_blob: Tensor) -> Optional[Tensor]:
    """field_backward + field_wgrad in one launch (csrc/field_bwd_fused.cu): chain CTAs and wgrad CTAs side by side."""
    g_area = torch.empty(n, s, device=stash.device, dtype=torch.float32) if want_area else None
    ws = torch.empty(_lib.lib().rsn_field_backward_fused_workspace_bytes(n * s), dtype=torch.uint8, device=stash.device)
    with _Prof("field_bwd_fused_kernel+area" if want_area else "field_bwd_fused_kernel", n * s):
        _lib.call("rsn_field_backward_fused", _lib.ptr(wblob_t), _lib.ptr(stash), mode, _lib.ptr(origins), _lib.ptr(dirs),
                  _lib.ptr(area), _lib.ptr(bins), n, s, _lib.ptr(g_sigma), _lib.ptr(g_feat), _lib.ptr(feat),
                  _lib.ptr(aux), _lib.ptr(dy_stash), _lib.ptr(g_area), _lib.ptr(grad_blob), _lib.ptr(ws), _lib.stream())
    return g_area


def field_wgrad(stash: Tensor, dy_stash: Tensor, n_points: int, grad_blob: Tensor) -> None:
    """K5 wgrad: accumulates dW / db of every Linear over the pass into grad_blob (fp32, wgrad_layout())."""
    with _Prof("field_wgrad_kernel", n_points):
        _lib.call("rsn_field_wgrad", _lib.ptr(stash), _lib.ptr(dy_stash), n_points, _lib.ptr(grad_blob), _lib.stream())


def wgrad_finish(grad_blob: Tensor, w_bott: Tensor, b_bott: Tensor, w_mid: Tensor) -> None:
    """Once per step, after the last field_wgrad (and the all-reduce): derive the bottleneck layer's gradients and the
    bottleneck columns of d mlp_mid.layers.0.weight from G = dY_mid^T h7 = rows 64-191 of wgrad job 10
    (include/rsn_b200.h: rsn_field_wgrad_finish).
    CPU tensors (the gloo test of the flush path) take the same algebra through torch."""
    if grad_blob.is_cuda:
        _lib.call("rsn_field_wgrad_finish", _lib.ptr(grad_blob), _lib.ptr(_f32c(w_bott.detach())),
                  _lib.ptr(_f32c(b_bott.detach())), _lib.ptr(_f32c(w_mid.detach())), _lib.stream())
        return
    offs, shapes, _ = wgrad_layout()
    g = grad_blob[offs[20] + 64 * 256: offs[20] + 192 * 256].view(128, 256)      # rows 64-191 of job 10's region
    db_mid = grad_blob[offs[21] + 64: offs[21] + 192]
    w_mb = w_mid.detach()[:, 34:]
    grad_blob[offs[18]: offs[18] + 256 * 256] = (w_mb.T @ g).reshape(-1)
    grad_blob[offs[19]: offs[19] + 256] = w_mb.T @ db_mid
    grad_blob[offs[24]: offs[24] + 128 * 256] = (g @ w_bott.detach().T + torch.outer(db_mid, b_bott.detach())).reshape(-1)
    grad_blob[offs[25]: offs[25] + 128] = db_mid


PACK_ORDER = ([f"mlp_base.layers.{l}.weight" for l in range(8)] + [f"mlp_base.layers.{l}.bias" for l in range(8)]
              + [f"{m}.{k}" for m in ("field_output_bottleneck.net", "mlp_mid.layers.0", "field_output_mid.net",
                                      "field_output_density.net", "field_output_normals.net",
                                      "field_output_roughness.net", "field_output_diff.net", "field_output_tint.net")
                 for k in ("weight", "bias")])


def pack_field(named_params) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """One-launch packing of the field's fp32 parameters (dict name -> CUDA tensor) into
    (forward blob, bias vector, transposed blob, bf16 density row) -- csrc/pack.cu."""
    tensors = [_f32c(named_params[k].detach()) for k in PACK_ORDER]
    dev = tensors[0].device
    ptrs = (_ct.c_void_p * len(tensors))(*[_lib.ptr(t) for t in tensors])
    wblob = torch.empty(_lib.lib().rsn_field_blob_bytes(), dtype=torch.uint8, device=dev)
    wblob_t = torch.empty(_lib.lib().rsn_field_blob_t_bytes(), dtype=torch.uint8, device=dev)
    bias = torch.empty(_lib.lib().rsn_field_bias_count(), dtype=torch.float32, device=dev)
    wd = torch.empty(256, dtype=torch.bfloat16, device=dev)
    _lib.call("rsn_pack_field", ptrs, _lib.ptr(wblob), _lib.ptr(wblob_t), _lib.ptr(bias), _lib.ptr(wd), _lib.stream())
    return wblob, bias, wblob_t, wd


@functools.lru_cache(maxsize=1)
def flat_layout():
    """(offsets of the 32 parameters of PACK_ORDER in the flat gradient vector, total length)."""
    offs = (_ct.c_int64 * 32)()
    total = _lib.lib().rsn_field_flat_layout(offs)
    return [int(o) for o in offs], int(total)


def unpack_grads_flat(grad_blob: Tensor, flat: Tensor) -> Tensor:
    """Gradient blob of field_wgrad -> flat fp32 gradient vector (one launch, csrc/pack.cu)."""
    _lib.call("rsn_unpack_grads", _lib.ptr(grad_blob), _lib.ptr(flat), _lib.stream())
    return flat


# ----------------------------------------------------------------------------------------- K9
def reflect_setup(comp16: Tensor, acc: Tensor, depth: Tensor, origins: Tensor, dirs: Tensor, clamp01: bool):
    """Per-ray quantities of the bounce from the composited fine pass (model.py:215-229,267-271), all detached:
    -> diff [N,3], tint [N,3], normal [N,3], n_dot_d [N,1], mask [N] bool, bounce origins [N,3], bounce dirs [N,3]."""
    comp16, acc, depth = _f32c(comp16.detach()), _f32c(acc.detach().reshape(-1)), _f32c(depth.detach().reshape(-1))
    origins, dirs = _f32c(origins), _f32c(dirs)
    n, dev = comp16.shape[0], comp16.device
    diff, tint, nrm, o2, wr = (torch.empty(n, 3, device=dev, dtype=torch.float32) for _ in range(5))
    ndd = torch.empty(n, 1, device=dev, dtype=torch.float32)
    mask = torch.empty(n, device=dev, dtype=torch.uint8)
    _lib.call("rsn_reflect_setup", _lib.ptr(comp16), _lib.ptr(acc), _lib.ptr(depth), _lib.ptr(origins), _lib.ptr(dirs),
              int(clamp01), _lib.ptr(diff), _lib.ptr(tint), _lib.ptr(nrm), _lib.ptr(ndd), _lib.ptr(mask), _lib.ptr(o2),
              _lib.ptr(wr), n, _lib.stream())
    return diff, tint, nrm, ndd, mask.bool(), o2, wr


class _ReflectCompose(torch.autograd.Function):
    """out = base; out[idx] = clip(diff[idx] + tint[idx] * (comp[:, :3] + bg * (1 - acc)), 0, 1)   (model.py:311-313)."""

    @staticmethod
    def forward(ctx, base, diff, tint, idx, comp, bg, acc, clamp_inner: bool):
        base, diff, tint, comp, bg, acc = (_f32c(t) for t in (base, diff, tint, comp, bg, acc.reshape(-1)))
        idx = idx.contiguous()
        n, m = base.shape[0], idx.shape[0]
        out = torch.empty_like(base)
        _lib.call("rsn_reflect_compose_fwd", _lib.ptr(base), _lib.ptr(diff), _lib.ptr(tint), _lib.ptr(idx), _lib.ptr(comp),
                  comp.shape[1], _lib.ptr(bg), _lib.ptr(acc), int(clamp_inner), _lib.ptr(out), n, m, _lib.stream())
        ctx.save_for_backward(diff, tint, idx, comp, bg, acc)
        return out

    @staticmethod
    def backward(ctx, g_out):
        diff, tint, idx, comp, bg, acc = ctx.saved_tensors
        g_out = _f32c(g_out)
        n, m = g_out.shape[0], idx.shape[0]
        g_base = torch.empty_like(g_out)
        g3 = torch.empty(m, 3, device=g_out.device, dtype=torch.float32)
        g_bg = torch.empty_like(g3)
        _lib.call("rsn_reflect_compose_bwd", _lib.ptr(g_out), _lib.ptr(diff), _lib.ptr(tint), _lib.ptr(idx), _lib.ptr(comp),
                  comp.shape[1], _lib.ptr(bg), _lib.ptr(acc), _lib.ptr(g3), _lib.ptr(g_bg), _lib.ptr(g_base), n, m,
                  _lib.stream())
        g_comp = torch.zeros_like(comp)
        g_comp[:, :3] = g3
        return g_base, None, None, None, g_comp, g_bg, None, None


def reflect_compose(base, diff, tint, idx, comp, bg, acc, clamp_inner: bool = False) -> Tensor:
    return _ReflectCompose.apply(base, diff, tint, idx, comp, bg, acc, clamp_inner)
