"""Training form of get_outputs: the same pass structure as the no-grad path (model.py), with each field pass an
autograd.Function over the hand-written forward / normals / dgrad / wgrad kernels, and the detach topology of
the reference expressed in the few per-ray torch ops between the kernels.

Gradient flow of one pass (csrc/field_bwd.cu, csrc/field_wgrad.cu):
    dL/d sigma, dL/d feat  --dgrad chain-->  dY of every Linear (bf16, HBM)  --wgrad-->  one fp32 gradient blob
The blob accumulates over all passes of a backward and is unpacked into the parameters' .grad ONCE, by an
autograd-engine callback queued from the first pass's backward; with data parallelism the blob is all-reduced
(one NCCL call of ~2.7 MB) right before the unpack -- the replacement of the reference's DDP wrapper
(reflect_sampling_nerf_pipeline.py:73-77).
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib, ops, packing


def _flush_grads(field) -> None:
    """Engine callback at the end of backward: (all-reduce and) unpack the gradient blob into .grad."""
    blob, field._grad_blob = field._grad_blob, None
    if blob is None:
        return
    if field.dp_world_size > 1 and not os.environ.get("RSN_DEBUG_SKIP_ALLREDUCE"):
        dist.all_reduce(blob, op=dist.ReduceOp.SUM)
        blob.mul_(1.0 / field.dp_world_size)          # DDP averages (pipeline.py:75)
    params = dict(field.named_parameters())
    # bottleneck layer: its gradients come from G = dY_mid^T h7 (linear in the blob, so after the all-reduce)
    ops.wgrad_finish(blob, params["field_output_bottleneck.net.weight"], params["field_output_bottleneck.net.bias"],
                     params["mlp_mid.layers.0.weight"])
    if blob.is_cuda:
        # one kernel: blob -> flat gradient vector; the parameters' .grad are views of it
        offs, total = ops.flat_layout()
        fresh = all(params[k].grad is None for k in ops.PACK_ORDER)
        flat = field.__dict__.get("_flat_grad")
        if flat is None or flat.device != blob.device or not fresh:
            flat = torch.empty(total, device=blob.device)      # (accumulating into existing grads: private buffer)
            if fresh:
                field.__dict__["_flat_grad"] = flat
        ops.unpack_grads_flat(blob, flat)
        for k, off in zip(ops.PACK_ORDER, offs):
            p = params[k]
            if not p.requires_grad:
                continue
            g = flat[off: off + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = g
            else:
                p.grad.add_(g)
        return
    offs, shapes, _ = ops.wgrad_layout()                    # host reference path (CPU tensors: the gloo test)
    grads = packing.unpack_grads(blob, offs, shapes)
    for name, p in params.items():
        g = grads.get(name)
        if g is None or not p.requires_grad:
            continue                                   # field_output_low never receives a gradient (App. B Q18)
        if p.grad is None:
            p.grad = g.clone(memory_format=torch.contiguous_format)
        else:
            p.grad.add_(g)


def _stash_workspace(field, n_points: int, device) -> Tensor:
    """The activation stash of the k-th field pass of a step lives in a persistent workspace (10.7 GB for a C2
    primary pass): a step's stashes are dead once its backward has run, and re-allocating ~28 GB per step makes
    the caching allocator thrash for the first several steps."""
    nbytes = _lib.lib().rsn_field_stash_bytes(n_points)
    pool = field.__dict__.setdefault("_stash_pool", [])
    k = field.__dict__.get("_stash_cursor", 0)
    field.__dict__["_stash_cursor"] = k + 1
    if k >= len(pool):
        pool.append(None)
    buf = pool[k]
    if buf is None or buf.numel() < nbytes or buf.device != device:
        pool[k] = None
        buf = torch.empty(int(nbytes * 1.25) if k >= 2 else nbytes, dtype=torch.uint8, device=device)   # reflected passes vary
        pool[k] = buf
    return buf


class _FieldPass(torch.autograd.Function):
    """One fused field evaluation over all samples of a ray batch (mode 0) or the infinity colour (mode 1)."""

    @staticmethod
    def forward(ctx, field, mode: int, primary: bool, origins, dirs, area, bins, *params):
        wblob, bias = field.packed()
        n_pts = dirs.shape[0] * (bins.shape[1] - 1 if bins is not None else 1)
        stash = _stash_workspace(field, n_pts, dirs.device)
        sigma, feat, stash, aux = ops.field_forward_train(wblob, bias, mode, origins, dirs, area, bins, stash)
        n, s = sigma.shape
        if primary:
            wblob_t, wd = field.packed_t()
            normals = ops.field_normals(wblob_t, wd, stash, n, s)     # K6; a constant (no create_graph upstream)
        else:
            normals = sigma.new_empty(0)
        ctx.field, ctx.mode = field, mode
        ctx.save_for_backward(origins if origins is not None else sigma.new_empty(0), dirs, area,
                              bins if bins is not None else sigma.new_empty(0), feat, aux, stash)
        ctx.mark_non_differentiable(normals)
        return sigma, feat, normals

    @staticmethod
    def backward(ctx, g_sigma, g_feat, _g_normals):
        field, mode = ctx.field, ctx.mode
        origins, dirs, area, bins, feat, aux, stash = ctx.saved_tensors
        n, s = feat.shape[0], feat.shape[1]
        none = (None,) * len(ctx.needs_input_grad)
        if g_sigma is None and g_feat is None:
            return none
        if g_feat is None:
            g_feat = torch.zeros_like(feat)
        want_area = bool(ctx.needs_input_grad[5])
        wblob_t, _ = field.packed_t()
        nbytes = _lib.lib().rsn_field_dy_stash_bytes(n * s)
        if field._dy_buffer is None or field._dy_buffer.numel() < nbytes or field._dy_buffer.device != feat.device:
            field._dy_buffer = torch.empty(nbytes, dtype=torch.uint8, device=feat.device)
        if field._grad_blob is None:
            field._grad_blob = torch.zeros(ops.wgrad_layout()[2], device=feat.device)
            torch.autograd.Variable._execution_engine.queue_callback(lambda: _flush_grads(field))
        args = (wblob_t, stash, mode, origins if mode == 0 else None, dirs, area.reshape(-1),
                bins if mode == 0 else None, n, s, None if g_sigma is None else g_sigma.contiguous(),
                g_feat.contiguous(), feat, aux, field._dy_buffer, want_area)
        if os.environ.get("RSN_FUSED_BWD", "0") == "1":
            # chain + wgrad CTAs in one launch (validated, opt-in): the wgrad's load rate is bound by the bytes one SM
            # can keep in flight (~45 GB/s per SM), so on half of the SMs it takes twice as long -- 8.2 ms fused against
            # 3.2 + 3.2 ms back to back at C2
            g_area = ops.field_backward_fused(*args, field._grad_blob)
        else:
            g_area = ops.field_backward(*args)
            ops.field_wgrad(stash, field._dy_buffer, n * s, field._grad_blob)
        out = list(none)
        if want_area:
            out[5] = g_area.sum(dim=1).reshape(area.shape)
        return tuple(out)


def field_pass(field, mode: int, primary: bool, origins, dirs, area, bins):
    return _FieldPass.apply(field, mode, primary, origins, dirs, area, bins, *[p for p in field.parameters()])


def _render(field, primary, o, d, area, eu_bins, detach_density: bool):
    sigma, feat, normals = field_pass(field, ops.MODE_SAMPLES, primary, o, d, area, eu_bins)
    if primary:   # the per-sample normal losses (model.py:403-407) ride on the compositing kernel
        w, acc, depth, comp, pnl, ol = ops.composite16(sigma, eu_bins, feat, normals)
        return feat, normals, w, acc[:, None], depth[:, None], comp, (pnl, ol)
    w, acc, depth, comp = ops.composite(sigma.detach() if detach_density else sigma, eu_bins, feat)
    return feat, normals, w, acc[:, None], depth[:, None], comp, None


def get_outputs_train(model, ray_bundle) -> Dict[str, Tensor]:
    """reflect_sampling_nerf_model.py:142-344 in training mode, with autograd."""
    field = model.field
    field.__dict__["_stash_cursor"] = 0          # stash workspaces are reused pass by pass, step after step
    o, d = ray_bundle.origins, ray_bundle.directions
    area, nears, fars = ray_bundle.pixel_area, ray_bundle.nears, ray_bundle.fars
    n, dev = o.shape[0], o.device
    clip01 = lambda x: torch.clip(x, 0.0, 1.0)  # noqa: E731

    # A. coarse (model.py:148-177)
    su, sp = model.sampler_uniform, model.sampler_pdf
    sp_c, eu_c = ops.sample_spaced(nears, fars, su.num_samples, su.kind, su.noise(n, dev))
    feat_c, nrm_c, w_c, acc_c, depth_c, comp_c, nl_c = _render(field, True, o, d, area, eu_c, False)
    rgb_c = clip01(comp_c[:, ops.F_RGB] + (1.0 - acc_c))
    # B. fine (model.py:182-211)
    sp_f, eu_f = ops.pdf_resample(w_c.detach(), sp_c, nears, fars, sp.num_samples, sp.kind, rand=sp.noise(n, dev),
                                  train=True)
    feat_f, nrm_f, w_f, acc_f, depth_f, comp_f, nl_f = _render(field, True, o, d, area, eu_f, False)
    rgb_f = clip01(comp_f[:, ops.F_RGB] + (1.0 - acc_f))
    # C. per-ray quantities of the bounce (model.py:215-229): everything detached except the roughness (one kernel)
    diff_r, tint_r, nrm_r, ndd, mask, o2_all, wr_all = ops.reflect_setup(comp_f, acc_f, depth_f, o, d, clamp01=False)
    rough = comp_f[:, ops.F_ROUGH_SIGMOID, None]                       # NOT detached (model.py:225-227)
    fallback = torch.ones(n, 3, device=dev) * (1.0 - acc_f)           # gradient to accumulation_fine (App. B Q10)
    outputs = {
        "mid_rgb_coarse": rgb_c, "mid_rgb_fine": rgb_f,
        "mid_reflect_coarse": fallback, "mid_reflect_fine": fallback,
        "accumulation_coarse": acc_c.detach(), "accumulation_fine": acc_f.detach(),
        "depth_coarse": depth_c, "depth_fine": depth_f,
        "weights_coarse": w_c.detach()[..., None], "weights_fine": w_f.detach()[..., None],
        "pred_normals_coarse": feat_c[..., ops.F_NORMAL], "pred_normals_fine": feat_f[..., ops.F_NORMAL],
        "normals_coarse": nrm_c, "normals_fine": nrm_f,
        "n_dot_d_coarse": feat_c[..., ops.F_NDOTD, None], "n_dot_d_fine": feat_f[..., ops.F_NDOTD, None],
        "diff": diff_r, "tint": tint_r, "roughness": rough, "mask": mask,
    }
    # fused per-ray sums of the normal losses, valid for exactly this outputs dict (get_loss_dict checks the identity)
    model.__dict__["_fused_normal_losses"] = (outputs["weights_fine"], {
        "predicted_normal_loss_coarse": nl_c[0], "orientation_loss_coarse": nl_c[1],
        "predicted_normal_loss_fine": nl_f[0], "orientation_loss_fine": nl_f[1]})
    idx = torch.nonzero(mask).reshape(-1)
    m = idx.numel()
    if m == 0:
        return outputs
    # D. reflected bundle (model.py:267-290): origins / directions detached, sqradius carries grad to the roughness
    o2, w_r = o2_all[idx], wr_all[idx]
    sqr = 2 * torch.abs(ndd[idx]) * rough[idx] ** 2
    area2 = math.pi * sqr
    nears2 = torch.zeros(m, 1, device=dev)
    fars2 = torch.full((m, 1), float(model.far), device=dev)
    _, feat_bg, _ = field_pass(field, ops.MODE_INF_COLOR, False, None, w_r, sqr, None)
    bg = feat_bg[:, 0, ops.F_RGB]
    # E. reflected coarse (model.py:292-313): weights detached => the reflected passes never train the density
    sr, sq = model.sampler_reciprocal, model.sampler_reflect_pdf
    sp_rc, eu_rc = ops.sample_spaced(nears2, fars2, sr.num_samples, sr.kind, sr.noise(m, dev))
    _, _, w_rc, acc_rc, _, comp_rc, _ = _render(field, False, o2, w_r, area2, eu_rc, True)
    outputs["mid_reflect_coarse"] = ops.reflect_compose(fallback, diff_r, tint_r, idx, comp_rc, bg, acc_rc.detach())
    # F. reflected fine (model.py:317-341)
    sp_rf, eu_rf = ops.pdf_resample(w_rc.detach(), sp_rc, nears2, fars2, sq.num_samples, sq.kind,
                                    rand=sq.noise(m, dev), train=True)
    _, _, w_rf, acc_rf, depth_rf, comp_rf, _ = _render(field, False, o2, w_r, area2, eu_rf, True)
    outputs["mid_reflect_fine"] = ops.reflect_compose(fallback, diff_r, tint_r, idx, comp_rf, bg, acc_rf.detach())
    outputs["depth_reflect_fine"] = depth_rf
    return outputs


class TrainStep:
    """One optimizer step of the hot path: get_outputs + get_loss_dict + backward (+ flat-gradient all-reduce)
    + RAdam (lr 1e-3, eps 1e-15: reflect_sampling_nerf_config.py:50-53).  Used by bench.py and the tests; under
    nerfstudio the Trainer does the same through the model's public methods."""

    def __init__(self, model, world_size: int = 1, lr: float = 1e-3) -> None:
        self.model = model
        model.field.dp_world_size = world_size
        self.opt = torch.optim.RAdam(model.get_param_groups()["fields"], lr=lr, eps=1e-15)

    def step(self, ray_bundle, image: Tensor) -> Tensor:
        self.opt.zero_grad(set_to_none=True)
        out = self.model(ray_bundle)
        loss_dict = self.model.get_loss_dict(out, {"image": image})
        loss = sum(loss_dict.values())
        loss.backward()
        self.opt.step()
        return loss.detach()
