"""get_outputs of the drop-in model (reflect_sampling_nerf_model.py:142-344) on the sm_100a kernels, for both the
no-grad form (eval / render) and the training form, in which every field pass is an autograd.Function over the
hand-written forward / normals / dgrad / wgrad kernels and the detach topology of the reference is
expressed by which tensors are handed to which Function.

The whole path is SYNC-FREE: the reference's boolean-mask indexing (model.py:229,267-289) reads the number of bouncing
rays M back to the host; here the mask is compacted on the device (ops.reflect_compact), the bounce passes are sized
for the capacity N and take M through a device counter, so a training step enqueues ~55 kernels without ever waiting
for the GPU and can be captured in a CUDA graph (TrainStep(graph=True)).

Gradient flow of one pass (csrc/field_bwd.cu, csrc/field_wgrad.cu):
    dL/d sigma, dL/d feat  --dgrad chain-->  dY of every Linear (bf16, HBM)  --wgrad-->  one fp32 gradient blob
The blob accumulates over all passes of a backward and is turned into the flat gradient vector ONCE, by an
autograd-engine callback queued from the first pass's backward; with data parallelism the blob is all-reduced
(one NCCL call of ~2.7 MB) right before -- the replacement of the reference's DDP wrapper
(reflect_sampling_nerf_pipeline.py:73-77).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib, ops, packing


# ----------------------------------------------------------------------------------------- gradient flush
def _allreduce_blob(field, blob) -> None:
    """Data parallelism: the ONE collective of a step -- mean of the gradient blob over the ranks (DDP averages,
    reflect_sampling_nerf_pipeline.py:73-77)."""
    timing = field.__dict__.get("_allreduce_events")
    if timing is not None and blob.is_cuda:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if blob.is_cuda and dist.get_backend() == "nccl":
        dist.all_reduce(blob, op=dist.ReduceOp.AVG)       # (NCCL averages inside the collective: no scaling launch)
    else:
        dist.all_reduce(blob, op=dist.ReduceOp.SUM)
        blob.mul_(1.0 / field.dp_world_size)
    if timing is not None and blob.is_cuda:
        e1.record()
        timing.append((e0, e1))


def _flush_grads(field, allreduce: bool = True) -> None:
    """Engine callback at the end of backward: (all-reduce and) unpack the gradient blob into .grad.
    allreduce=False: the caller has already reduced the blob (TrainStep's split CUDA graphs)."""
    blob, field._grad_blob = field._grad_blob, None
    if blob is None:
        return
    if field.dp_world_size > 1 and allreduce:
        _allreduce_blob(field, blob)
    params = dict(field.named_parameters())
    # bottleneck layer: its gradients come from G = dY_mid^T h7 (linear in the blob, so after the all-reduce)
    ops.wgrad_finish(blob, params["field_output_bottleneck.net.weight"], params["field_output_bottleneck.net.bias"],
                     params["mlp_mid.layers.0.weight"])
    if blob.is_cuda:
        # one kernel: blob -> flat gradient vector; the parameters' .grad are views of it
        offs, total = ops.flat_layout()
        fresh = all(params[k].grad is None for k in ops.PACK_ORDER)
        flat = field.__dict__.get("_flat_grad")
        views_ok = flat is not None and flat.device == blob.device and all(
            params[k].grad is not None and params[k].grad.data_ptr() == flat.data_ptr() + 4 * off
            for k, off in zip(ops.PACK_ORDER, offs))
        if views_ok and field.__dict__.get("_flat_grad_overwrite", False):
            ops.unpack_grads_flat(blob, flat)              # TrainStep: .grad already are the views; overwrite in place
            return
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


# ----------------------------------------------------------------------------------------- stash slots
class _StashSlot:
    """One persistent activation-stash workspace (10.7 GB for a C2 primary pass).  A step's stashes are dead once its
    backward has run, and re-allocating ~28 GB per step makes the caching allocator thrash, so the buffers are reused --
    but a slot belongs to ONE autograd graph at a time: `in_flight` from the forward that filled it until the backward
    that consumed it, and `generation` lets a late backward (retain_graph, two forwards before one backward) detect that
    its slot was handed to another forward instead of silently reading foreign activations."""
    __slots__ = ("buf", "in_flight", "generation")

    def __init__(self, buf: Tensor) -> None:
        self.buf, self.in_flight, self.generation = buf, False, 0


def _claim_stash(field, n_points: int, device) -> _StashSlot:
    nbytes = _lib.lib().rsn_field_stash_bytes(n_points)
    pool = field.__dict__.setdefault("_stash_pool", [])
    best = None
    for slot in pool:
        if not slot.in_flight and slot.buf.device == device and slot.buf.numel() >= nbytes:
            if best is None or slot.buf.numel() < best.buf.numel():
                best = slot
    if best is None:
        for i, slot in enumerate(pool):        # grow a free slot that is too small rather than keeping both
            if not slot.in_flight and slot.buf.device == device:
                pool.pop(i)
                break
        best = _StashSlot(torch.empty(nbytes, dtype=torch.uint8, device=device))
        pool.append(best)
    best.in_flight = True
    best.generation += 1
    return best


# ----------------------------------------------------------------------------------------- one field pass
class _FieldPass(torch.autograd.Function):
    """One fused field evaluation over all samples of a ray batch (mode 0) or the infinity colour (mode 1)."""
    always_backward = True        # (ops.Tape) the backward also releases the stash slot

    @staticmethod
    def forward(ctx, field, mode: int, primary: bool, count: Optional[Tensor], origins, dirs, area, bins, *params):
        wblob, bias = field.packed()
        n_pts = dirs.shape[0] * (bins.shape[1] - 1 if bins is not None else 1)
        slot = _claim_stash(field, n_pts, dirs.device)
        sigma, feat, stash, aux = ops.field_forward_train(wblob, bias, mode, origins, dirs, area, bins, slot.buf, count)
        n, s = sigma.shape
        if primary:
            wblob_t, wd = field.packed_t()
            normals = ops.field_normals(wblob_t, wd, stash, n, s)     # K6; a constant (no create_graph upstream)
        else:
            normals = sigma.new_empty(0)
        ctx.field, ctx.mode, ctx.count, ctx.slot, ctx.slot_generation = field, mode, count, slot, slot.generation
        ctx.save_for_backward(origins if origins is not None else sigma.new_empty(0), dirs, area,
                              bins if bins is not None else sigma.new_empty(0), feat, aux)
        ctx.mark_non_differentiable(normals)
        ctx.set_materialize_grads(False)
        return sigma, feat, normals

    @staticmethod
    def backward(ctx, g_sigma, g_feat, _g_normals):
        field, mode, slot = ctx.field, ctx.mode, ctx.slot
        origins, dirs, area, bins, feat, aux = ctx.saved_tensors
        n, s = feat.shape[0], feat.shape[1]
        none = (None,) * len(ctx.needs_input_grad)
        if g_sigma is None and g_feat is None:
            slot.in_flight = False
            return none
        if slot.generation != ctx.slot_generation:
            raise RuntimeError(
                "rsn_b200: the activation stash of this field pass was handed to a later forward pass after a first "
                "backward released it (retain_graph=True / a second backward over the same graph is not supported "
                "once another training forward has run)")
        if g_feat is None:
            g_feat = torch.zeros_like(feat)
        want_area = bool(ctx.needs_input_grad[6])
        wblob_t, _ = field.packed_t()
        nbytes = _lib.lib().rsn_field_dy_stash_bytes(n * s)
        if field._dy_buffer is None or field._dy_buffer.numel() < nbytes or field._dy_buffer.device != feat.device:
            field._dy_buffer = torch.empty(nbytes, dtype=torch.uint8, device=feat.device)
        if field._grad_blob is None:
            blob = field.__dict__.get("_grad_blob_static")
            if blob is None or blob.device != feat.device:
                blob = torch.empty(ops.wgrad_layout()[2], device=feat.device)
                field.__dict__["_grad_blob_static"] = blob
            blob.zero_()
            field._grad_blob = blob
            if ops.TAPE is None:          # (the tape's owner flushes explicitly)
                torch.autograd.Variable._execution_engine.queue_callback(lambda: _flush_grads(field))
        g_area = ops.field_backward(wblob_t, slot.buf, mode, origins if mode == 0 else None, dirs, area.reshape(-1),
                                    bins if mode == 0 else None, n, s, None if g_sigma is None else g_sigma.contiguous(),
                                    g_feat.contiguous(), feat, aux, field._dy_buffer, want_area, ctx.count)
        ops.field_wgrad(slot.buf, field._dy_buffer, n * s, field._grad_blob, ctx.count, s)
        slot.in_flight = False            # stream-ordered: the next forward's writes queue behind this wgrad
        out = list(none)
        if want_area:
            out[6] = g_area.sum(dim=1).reshape(area.shape)
        return tuple(out)


def field_pass(field, mode: int, primary: bool, origins, dirs, area, bins, count: Optional[Tensor] = None):
    return ops.tape_apply(_FieldPass, field, mode, primary, count, origins, dirs, area, bins,
                          *[p for p in field.parameters()])


# ----------------------------------------------------------------------------------------- the path
def _render(field, grad: bool, primary: bool, o, d, area, eu_bins, count=None):
    """sampled bins -> fused field -> compositing (model.py:151-177 and its three repeats).
    -> feat, normals, weights, acc [N], depth [N], comp16, (pnl, ol) or None, rgb_blend or None"""
    if grad:
        sigma, feat, normals = field_pass(field, ops.MODE_SAMPLES, primary, o, d, area, eu_bins, count)
    else:
        sigma, feat = field.evaluate_samples(o, d, area, eu_bins, count)
        normals = None
    if primary:   # the per-sample normal losses (model.py:403-407) and the white blend + clip ride on the compositing kernel
        w, acc, depth, comp, pnl, ol, rgb = ops.composite16(sigma, eu_bins, feat, normals if grad else None, None,
                                                            blend=True)
        return feat, normals, w, acc, depth, comp, ((pnl, ol) if grad else None), rgb
    # bounce passes: weights detached => they never train the density (model.py:297,323)
    w, acc, depth, comp, _, _, _ = ops.composite16(sigma, eu_bins, feat, None, count, blend=False, detach_sigma=True)
    return feat, normals, w, acc, depth, comp, None, None


def _stratification_noise(samplers, n: int, dev):
    """The four samplers' stratification noise of one step.  Training without injected noise: ONE torch.rand launch whose
    contiguous chunks the samplers share (instead of one launch per sampler); otherwise each sampler's own noise()."""
    plain = all(s.train_stratified and s.training and s.injected_rand is None and not getattr(s, "single_jitter", False)
                for s in samplers)
    if not plain:
        return [s.noise(n, dev) for s in samplers]
    cols = [s.num_samples + 1 for s in samplers]
    flat = torch.rand(n * sum(cols), device=dev)
    out, off = [], 0
    for c in cols:
        out.append(flat[off: off + n * c].view(n, c))
        off += n * c
    return out


def get_outputs(model, ray_bundle, grad: bool) -> Dict[str, Tensor]:
    """reflect_sampling_nerf_model.py:142-344.  grad=True: training mode with autograd."""
    field = model.field
    o, d = ray_bundle.origins, ray_bundle.directions
    area, nears, fars = ray_bundle.pixel_area, ray_bundle.nears, ray_bundle.fars
    n, dev = o.shape[0], o.device
    training = model.training

    # A. coarse (model.py:148-177)
    su, sp = model.sampler_uniform, model.sampler_pdf
    sr, sq = model.sampler_reciprocal, model.sampler_reflect_pdf
    noise_u, noise_p, noise_r, noise_q = _stratification_noise((su, sp, sr, sq), n, dev)
    sp_c, eu_c = ops.sample_spaced(nears, fars, su.num_samples, su.kind, noise_u)
    feat_c, nrm_c, w_c, acc_c, depth_c, comp_c, nl_c, rgb_c = _render(field, grad, True, o, d, area, eu_c)
    # B. fine (model.py:182-211)
    sp_f, eu_f = ops.pdf_resample(w_c.detach(), sp_c, nears, fars, sp.num_samples, sp.kind, rand=noise_p,
                                  train=training)
    feat_f, nrm_f, w_f, acc_f, depth_f, comp_f, nl_f, rgb_f = _render(field, grad, True, o, d, area, eu_f)
    # C. per-ray quantities of the bounce (model.py:215-229): everything detached except the roughness (one kernel),
    #    then the device-side compaction of the mask
    diff_r, tint_r, nrm_r, ndd, mask, o2_all, wr_all = ops.reflect_setup(comp_f, acc_f, depth_f, o, d,
                                                                         clamp01=not training)
    idx, inv, count = ops.reflect_compact(mask)
    rough = comp_f[:, ops.F_ROUGH_SIGMOID, None]                       # NOT detached (model.py:225-227)
    if not grad:
        nrm_c, nrm_f = feat_c[..., ops.F_NORMAL], feat_f[..., ops.F_NORMAL]     # eval: normals = predicted normals (Q8)
    outputs = {
        "mid_rgb_coarse": rgb_c, "mid_rgb_fine": rgb_f,
        "accumulation_coarse": acc_c.detach()[:, None], "accumulation_fine": acc_f.detach()[:, None],
        "depth_coarse": depth_c[:, None], "depth_fine": depth_f[:, None],
        "weights_coarse": w_c.detach()[..., None], "weights_fine": w_f.detach()[..., None],
        "pred_normals_coarse": feat_c[..., ops.F_NORMAL], "pred_normals_fine": feat_f[..., ops.F_NORMAL],
        "normals_coarse": nrm_c, "normals_fine": nrm_f,
        "n_dot_d_coarse": feat_c[..., ops.F_NDOTD, None], "n_dot_d_fine": feat_f[..., ops.F_NDOTD, None],
        "diff": diff_r, "tint": tint_r, "roughness": rough, "mask": mask.view(torch.bool),
    }
    if grad:
        # fused per-ray sums of the normal losses, valid for exactly this outputs dict (get_loss_dict checks the identity)
        model.__dict__["_fused_normal_losses"] = (outputs["weights_fine"], {
            "predicted_normal_loss_coarse": nl_c[0], "orientation_loss_coarse": nl_c[1],
            "predicted_normal_loss_fine": nl_f[0], "orientation_loss_fine": nl_f[1]})
    # D. reflected bundle (model.py:267-290): origins / directions detached, sqradius carries grad to the roughness;
    #    rows >= count of everything below are never touched
    o2, w_r, sqr, area2 = ops.reflect_bundle(comp_f if grad else comp_f.detach(), idx, inv, count, o2_all, wr_all, ndd)
    nears2, fars2 = model._bounce_planes(n, dev)                       # zeros * near (App. B Q4), ones * far
    if grad:
        _, feat_bg, _ = field_pass(field, ops.MODE_INF_COLOR, False, None, w_r, sqr, None, count)
        bg = ops.inf_color_rgb(feat_bg)
    else:
        bg = field.get_inf_color(w_r, sqr, count)
    # E. reflected coarse (model.py:292-313)
    sp_rc, eu_rc = ops.sample_spaced(nears2, fars2, sr.num_samples, sr.kind, noise_r, count)
    _, _, w_rc, acc_rc, _, comp_rc, _, _ = _render(field, grad, False, o2, w_r, area2, eu_rc, count)
    outputs["mid_reflect_coarse"], _ = ops.reflect_compose(acc_f, diff_r, tint_r, inv, comp_rc, bg, acc_rc.detach(),
                                                           None, clamp_inner=not training)
    # F. reflected fine (model.py:317-341)
    sp_rf, eu_rf = ops.pdf_resample(w_rc.detach(), sp_rc, nears2, fars2, sq.num_samples, sq.kind,
                                    rand=noise_q, train=training, count=count)
    _, _, w_rf, acc_rf, depth_rf, comp_rf, _, _ = _render(field, grad, False, o2, w_r, area2, eu_rf, count)
    outputs["mid_reflect_fine"], depth_pad = ops.reflect_compose(acc_f, diff_r, tint_r, inv, comp_rf, bg,
                                                                 acc_rf.detach(), depth_rf, clamp_inner=not training)
    # [N,1], 0 where the ray did not bounce: the reference's ragged [M,1] (model.py:341) breaks upstream's full-image
    # assembly (App. B Q13); outputs["depth_reflect_fine"][outputs["mask"]] is the reference's tensor
    outputs["depth_reflect_fine"] = depth_pad[:, None]
    model.__dict__["last_num_bounced"] = count                         # device int32 [1]: M, never read by the path itself
    return outputs


def get_outputs_train(model, ray_bundle) -> Dict[str, Tensor]:
    return get_outputs(model, ray_bundle, True)


# ----------------------------------------------------------------------------------------- one optimizer step
def sync_parameters(field, src: int = 0) -> None:
    """Data parallel start-up: every rank takes rank `src`'s parameters.  The reference's DistributedDataParallel wrapper
    (reflect_sampling_nerf_pipeline.py:73-77) broadcasts them at construction; nerfstudio seeds each rank differently
    (machine.seed + global_rank), so without this the replicas would start from -- and keep -- different weights."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    with torch.no_grad():
        for p in field.parameters():
            dist.broadcast(p.data, src=src)
        for b in field.buffers():
            dist.broadcast(b.data, src=src)
    field._packed = None          # derived bf16 operand blobs follow the new values


class TrainStep:
    """One optimizer step of the hot path: get_outputs + get_loss_dict + backward (+ flat-gradient all-reduce)
    + fused RAdam / exponential decay / bf16 re-pack (reflect_sampling_nerf_config.py:50-53).  Used by bench.py and the
    tests; under nerfstudio the Trainer does the same through the model's public methods, torch.autograd and
    reflect_sampling_nerf_b200.optim.FusedRAdam.

    The backward is replayed from ops.Tape (the same Function.backward kernels torch.autograd would call, in the same
    order, on one thread and one stream); autograd=True runs it through torch.autograd instead (A/B tests).
    graph=True captures the whole step (forward, backward, optimizer) in a CUDA graph after 3 eager steps; inputs are
    copied into static buffers and the step is one cudaGraphLaunch.  With world_size > 1 the step is TWO graphs around the
    one collective -- [forward + losses + backward] -> eager NCCL all-reduce of the gradient blob -> [finish + unpack +
    RAdam + re-pack] -- so that NCCL never runs inside a capture."""

    def __init__(self, model, world_size: int = 1, lr: float = 1e-3, lr_final: float = 0.0, max_steps: int = 0,
                 graph: bool = False, torch_optimizer: bool = False, autograd: bool = False) -> None:
        from .optim import FusedRAdam
        self.model = model
        model.field.dp_world_size = world_size
        if world_size > 1:
            sync_parameters(model.field)
        params = model.get_param_groups()["fields"]
        if torch_optimizer:      # the reference's optimizer, for A/B tests
            self.opt = torch.optim.RAdam(params, lr=lr, eps=1e-15)
        else:
            self.opt = FusedRAdam(params, lr=lr, eps=1e-15, lr_final=lr_final, max_steps=max_steps, field=model.field)
        self.fused = not torch_optimizer
        self.autograd = autograd
        self.graph_requested, self.graph, self.static = graph and self.fused and not autograd, None, None
        self.steps_done = 0
        self._one = None

    def _forward_backward_tape(self, ray_bundle, image: Tensor, flush: bool = True) -> Tensor:
        """flush=False leaves the un-reduced gradient blob on the field (field._grad_blob) for finish_step()."""
        field = self.model.field
        if not self.model.training:
            raise RuntimeError("TrainStep needs the model in training mode")
        tape = ops.Tape()
        for p in field.parameters():
            tape.live.add(id(p))
        with torch.no_grad():
            ops.TAPE = tape
            try:
                out = self.model(ray_bundle)
                self.model.get_loss_dict(out, {"image": image})
                total = self.model.__dict__.pop("_fused_loss_total")
                if self._one is None or self._one.device != total.device:
                    self._one = torch.ones((), device=total.device)
                tape.backward(total, self._one)
            finally:
                ops.TAPE = None
                tape.records.clear()
            if flush:
                _flush_grads(field)
        self.last_outputs = out
        return total

    def allreduce_grads(self) -> None:
        """The step's one collective, on the blob _forward_backward_tape(flush=False) left behind (no-op at world_size 1)."""
        field = self.model.field
        if field.dp_world_size > 1 and field._grad_blob is not None:
            _allreduce_blob(field, field._grad_blob)

    def finish_step(self) -> None:
        """Second half of a split step: reduced blob -> flat gradient vector -> fused RAdam + re-pack."""
        with torch.no_grad():
            _flush_grads(self.model.field, allreduce=False)
        self.opt.step()

    def _eager(self, ray_bundle, image: Tensor) -> Tensor:
        if self.autograd or not self.fused:
            if not self.fused:
                self.opt.zero_grad(set_to_none=True)
            out = self.model(ray_bundle)
            loss_dict = self.model.get_loss_dict(out, {"image": image})
            total = self.model.__dict__.pop("_fused_loss_total", None)
            loss = total if total is not None else sum(loss_dict.values())
            loss.backward()
            self.last_outputs = out
            loss = loss.detach()
        else:
            loss = self._forward_backward_tape(ray_bundle, image)
        self.opt.step()
        return loss

    def step(self, ray_bundle, image: Tensor) -> Tensor:
        """ray_bundle / image may live on the device or in (pinned) HOST memory -- the form a CPU data loader hands over:
        host batches are copied to the device here; once the step is a CUDA graph, straight into its static input buffers."""
        self.steps_done += 1
        if not ray_bundle.origins.is_cuda and not (self.graph_requested and self.graph is not None):
            from .rays import RayBundle
            dev = next(self.model.field.parameters()).device
            ray_bundle = RayBundle(origins=ray_bundle.origins.to(dev, non_blocking=True),
                                   directions=ray_bundle.directions.to(dev, non_blocking=True),
                                   pixel_area=ray_bundle.pixel_area.to(dev, non_blocking=True))
            image = image.to(dev, non_blocking=True)
        if not self.graph_requested:
            return self._eager(ray_bundle, image)
        if self.graph is None:
            if self.steps_done <= 3:          # eager warm-up: workspaces, caches and function attributes settle
                return self._eager(ray_bundle, image)
            from .rays import RayBundle
            st = {"o": ray_bundle.origins.clone(), "d": ray_bundle.directions.clone(),
                  "a": ray_bundle.pixel_area.clone(), "img": image.clone()}
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            split = self.model.field.dp_world_size > 1
            with torch.cuda.graph(g):
                bundle = RayBundle(origins=st["o"], directions=st["d"], pixel_area=st["a"])
                if split:
                    st["loss"] = self._forward_backward_tape(bundle, st["img"], flush=False)
                else:
                    st["loss"] = self._eager(bundle, st["img"])
            if split:                          # (the blob is a persistent buffer of the field: both graphs address it)
                st["blob"] = self.model.field._grad_blob
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, pool=g.pool()):
                    self.finish_step()
                st["finish"] = g2
            self.graph, self.static = g, st
            return self._replay()              # capture does not execute: run the step it recorded on this batch
        st = self.static
        st["o"].copy_(ray_bundle.origins, non_blocking=True)
        st["d"].copy_(ray_bundle.directions, non_blocking=True)
        st["a"].copy_(ray_bundle.pixel_area, non_blocking=True)
        st["img"].copy_(image, non_blocking=True)
        return self._replay()

    def _replay(self) -> Tensor:
        st = self.static
        self.graph.replay()
        if "finish" in st:
            _allreduce_blob(self.model.field, st["blob"])
            st["finish"].replay()
        return st["loss"]
