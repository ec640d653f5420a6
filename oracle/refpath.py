"""ORACLE (test infrastructure): staged CPU/PyTorch-fp32 restatement of the reference's
per-ray hot path.  PINNED against the unmodified reference files executed on
`oracle.nerfstudio_shim` (tests/golden/make_golden.py -> tests/test_oracle_golden.py).

Reference anchors (paths relative to /root/reference/reflect_sampling_nerf/):
  field:       reflect_sampling_nerf_field.py:36-207
  components:  reflect_sampling_nerf_components.py:14-140
  get_outputs: reflect_sampling_nerf_model.py:142-344
  losses:      reflect_sampling_nerf_model.py:346-430
  warm-up:     reflect_sampling_nerf_pipeline.py:79-91

The restatement is organised by *stage* (the K1..K9 split of SURVEY.md §2.4) rather than
by the reference's method layout, so each CUDA kernel has a stage oracle of its own:

  spaced_bins / pdf_bins          K1 / K2   (bit-exact targets)
  frustum_gaussian / contract     K3
  ipe / ide                       K4 / K7
  OracleField.point_heads         K5 (+K6 density-gradient normals)
  composite                       K8
  reflect_setup                   K9
  get_outputs / get_loss_dict     the whole path, same detach topology

Quirks marked "preserve" in SURVEY.md Appendix B are kept and tagged Q<n>.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import upstream as U

NEAR_REFLECT = 1.0 / 16   # model.py:114 (multiplied into zeros -> 0, Q4)
FAR_REFLECT = 2.0 ** 8    # model.py:113
RECIPROCAL_TAN = 0.25     # model.py:111

LOSS_COEFFICIENTS = {     # model.py:56-69
    "loss_low_coarse": 1e-1, "loss_low_fine": 1e-1,
    "loss_mid_coarse": 1.0, "loss_mid_fine": 1.0,
    "loss_reflect_low_coarse": 1e-1, "loss_reflect_low_fine": 1e-1,
    "loss_reflect_mid_coarse": 1.0, "loss_reflect_mid_fine": 1.0,
    "predicted_normal_loss_coarse": 3e-5, "predicted_normal_loss_fine": 3e-4,
    "orientation_loss_coarse": 1e-2, "orientation_loss_fine": 1e-1,
}


def warmup_coefficients(step: int, coeffs: Dict[str, float]) -> Dict[str, float]:
    """pipeline.py:79-91 (Q14): normal/orientation terms are off for step < 50."""
    on = step >= 50
    coeffs["predicted_normal_loss_coarse"] = 3e-5 if on else 0.0
    coeffs["predicted_normal_loss_fine"] = 3e-4 if on else 0.0
    coeffs["orientation_loss_coarse"] = 1e-2 if on else 0.0
    coeffs["orientation_loss_fine"] = 1e-1 if on else 0.0
    return coeffs


# ----------------------------------------------------------------------------- K1 / K2
def spacing_fns(kind: str):
    """kind = 'uniform' (model.py:109) or 'reciprocal' (components.py:32-33, tan=0.25)."""
    if kind == "uniform":
        return (lambda x: x), (lambda x: x)
    if kind == "reciprocal":
        tan = RECIPROCAL_TAN
        return (lambda x: x / (1 / tan + x)), (lambda x: x / tan / (1 - x))
    raise ValueError(kind)


def make_spaced_sampler(kind: str, num_samples: int) -> U.SpacedSampler:
    fn, inv = spacing_fns(kind)
    return U.SpacedSampler(spacing_fn=fn, spacing_fn_inv=inv, num_samples=num_samples)


def spaced_bins(nears: Tensor, fars: Tensor, num_samples: int, kind: str,
                t_rand: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """K1 stage oracle -> (spacing bins [N,S+1], euclidean bins [N,S+1]).  `t_rand` None = eval."""
    fn, inv = spacing_fns(kind)
    bins = torch.linspace(0.0, 1.0, num_samples + 1)[None, :]
    if t_rand is not None:
        c = (bins[..., 1:] + bins[..., :-1]) / 2.0
        upper = torch.cat([c, bins[..., -1:]], -1)
        lower = torch.cat([bins[..., :1], c], -1)
        bins = lower + (upper - lower) * t_rand
    s_near, s_far = fn(nears), fn(fars)
    euclid = inv(bins * s_far + (1 - bins) * s_near)
    return bins.expand_as(euclid), euclid


def pdf_bins(weights: Tensor, spacing_bins: Tensor, nears: Tensor, fars: Tensor, kind: str,
             num_samples: int, rand: Optional[Tensor]) -> Tuple[Tensor, Tensor, Tensor]:
    """K2 stage oracle -> (new spacing bins, new euclidean bins, searchsorted indices).
    weights [N,S]; spacing_bins [N,S+1]; rand [N,num_samples+1] in [0,1) or None (eval)."""
    fn, inv = spacing_fns(kind)
    sampler = U.PDFSampler(num_samples=num_samples, include_original=False)
    sampler.train(rand is not None)
    sampler.injected_rand = rand
    n = weights.shape[0]
    s_near, s_far = fn(nears), fn(fars)
    to_euclid = lambda x: inv(x * s_far + (1 - x) * s_near)  # noqa: E731
    dummy = torch.zeros(n, 3)
    bundle = U.RayBundle(origins=dummy, directions=dummy, pixel_area=torch.zeros(n, 1), nears=nears, fars=fars)
    rs = U.RaySamples(frustums=None, spacing_starts=spacing_bins[:, :-1, None],
                      spacing_ends=spacing_bins[:, 1:, None], spacing_to_euclidean_fn=to_euclid)
    out = sampler(bundle, rs, weights[..., None])
    new_spacing = torch.cat([out.spacing_starts[..., 0], out.spacing_ends[:, -1:, 0]], -1)
    new_euclid = torch.cat([out.frustums.starts[..., 0], out.frustums.ends[:, -1:, 0]], -1)
    return new_spacing, new_euclid, sampler.last_inds


# ----------------------------------------------------------------------------- K3
def frustum_gaussian(origins, directions, starts, ends, pixel_area):
    """field.py:90-96 -> Frustums.get_gaussian_blob (App. A.1)."""
    fr = U.Frustums(origins=origins, directions=directions, starts=starts, ends=ends, pixel_area=pixel_area)
    g = fr.get_gaussian_blob()
    return g.mean, g.cov


def contract(mean: Tensor, cov: Tensor) -> Tuple[Tensor, Tensor]:
    """field.py:98-119: mip-NeRF-360 contraction of the mean and J cov J with the closed-form
    (symmetric) Jacobian; diagonal clamped at 0 (Q9)."""
    n2 = torch.sum(mean**2, dim=-1, keepdim=True)
    n1 = torch.sqrt(n2)
    outside = n1 > 1
    mean_c = torch.where(outside, (2 * n1 - 1) / n2 * mean, mean)
    n1e, n2e = n1.unsqueeze(-1), n2.unsqueeze(-1)
    proj = mean[..., :, None] * mean[..., None, :] / n2e
    eye = torch.eye(3, device=mean.device).expand(proj.shape)
    jac = torch.where(outside[..., None], ((2 * n1e - 2) * (eye - proj) + eye) / n2e, eye)
    cov_c = torch.matmul(torch.matmul(jac, cov), jac)
    for i in range(3):
        cov_c[..., i, i] = F.relu(cov_c[..., i, i])
    return mean_c, cov_c


# ----------------------------------------------------------------------------- K4 / K7
_IPE = U.NeRFEncoding(in_dim=3, num_frequencies=16, min_freq_exp=0.0, max_freq_exp=16.0, include_input=True)


def ipe(mean: Tensor, cov: Tensor) -> Tensor:
    """model.py:98-100 + field.py:129 -> [...,99] (App. A.4)."""
    return _IPE(mean, covs=cov)


IDE_DECAY = (1.0, 3.0, 10.0, 36.0)   # l(l+1)/2 for l = 1,2,4,8 (components.py:136-139)
IDE_BANDS = ((0, 3), (3, 8), (8, 17), (17, 34))


@torch.no_grad()
def ide_basis(d: Tensor) -> Tensor:
    """components.py:52-132: the 34 hand-expanded polynomials, constants verbatim (Q3: entries 18
    and 32 carry 5.8314..., kept)."""
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    out = torch.zeros((*d.shape[:-1], 34), device=d.device)
    x2, y2, z2 = x**2, y**2, z**2
    xy, xz, yz = x * y, x * z, y * z
    dxy = x2 - y2
    a = 3 * x2 - y2          # "y2_3x2"
    b = x2 - 3 * y2          # "x2_3y2"
    z4 = z**4
    x4, y4 = x**4, y**4
    p = y4 - 10 * x2 * y2 + 5 * x4
    q = x4 - 10 * x2 * y2 + 5 * y4
    r = (x2 - 5 * y2) * 7 * x4 + (21 * x2 - y2) * y4
    s = (x2 - 21 * y2) * x4 + (5 * x2 - y2) * 7 * y4
    c1 = 0.48860251190291992
    terms = [
        c1 * y, c1 * z, c1 * x,
        1.09254843059207907 * xy,
        1.09254843059207907 * yz,
        0.31539156525252001 * (3 * z2 - 1),
        1.09254843059207907 * xz,
        0.54627421529603953 * dxy,
        2.50334294179670453 * xy * dxy,
        1.77013076977993053 * yz * a,
        0.94617469575756001 * xy * (7 * z2 - 1),
        0.66904654355728916 * yz * (7 * z2 - 3),
        0.1057855469152043038 * (35 * z4 - 30 * z2 + 3),
        0.66904654355728916 * xz * (7 * z2 - 3),
        0.473087347878780009 * dxy * (7 * z2 - 1),
        1.77013076977993053 * xz * b,
        0.62583573544917613 * (x2 * b - y2 * a),
        5.83141328139863895 * xy * (x2 * x4 - 7 * x4 * y2 + 7 * x2 * y4 - y2 * y4),
        5.83141328139863895 * yz * r,
        1.06466553211908514 * xy * (15 * z2 - 1) * (3 * x4 - 10 * x2 * y2 + 3 * y4),
        3.44991062209810801 * yz * (5 * z2 - 1) * p,
        1.91366609903732278 * xy * (65 * z4 - 26 * z2 + 1) * dxy,
        1.23526615529554407 * yz * (39 * z4 - 26 * z2 + 3) * a,
        0.91230451686981894 * xy * (143 * z4 * z2 - 143 * z4 + 33 * z2 - 1),
        0.1090412458987799555 * yz * (715 * z4 * z2 - 1001 * z4 + 385 * z2 - 35),
        0.0090867704915649962938 * (6435 * z4 * z4 - 12012 * z4 * z2 + 6930 * z4 - 1260 * z2 + 35),
        0.1090412458987799555 * xz * (715 * z4 * z2 - 1001 * z4 + 385 * z2 - 35),
        0.456152258434909470 * (143 * z4 * z2 - 143 * z4 + 33 * z2 - 1) * dxy,
        1.23526615529554407 * xz * (39 * z4 - 26 * z2 + 3) * b,
        0.478416524759330697 * (65 * z4 - 26 * z2 + 1) * (x2 * b - y2 * a),
        3.44991062209810801 * xz * (5 * z2 - 1) * q,
        0.53233276605954257 * (15 * z2 - 1) * (x2 * q - y2 * p),
        5.83141328139863895 * xz * s,
        0.72892666017482986 * (x2 * s - y2 * r),
    ]
    for i, t in enumerate(terms):
        out[..., i] = t
    return out


@torch.no_grad()
def ide(d: Tensor, roughness: Tensor) -> Tensor:
    """components.py:134-140.  The whole encoder runs under no_grad (components.py:52) and is fed
    the VIEW direction (Q1) with a detached Softplus roughness (model.py:174, Q2)."""
    out = ide_basis(d)
    for (lo, hi), k in zip(IDE_BANDS, IDE_DECAY):
        out[..., lo:hi] *= torch.exp(-roughness * k)
    return out


# ----------------------------------------------------------------------------- K5 / K6
class OracleField(U.Field):
    """Same sub-module names, shapes and construction order as field.py:36-86, so state dicts and
    seeded random initialisations are interchangeable with the reference field."""

    def __init__(self, density_bias: float = 0.5) -> None:
        super().__init__()
        self.position_encoding = _IPE
        self.mlp_base = U.MLP(in_dim=99, num_layers=8, layer_width=256, skip_connections=(4,),
                              out_activation=nn.ReLU())
        self.field_output_density = U.DensityFieldHead(in_dim=256, activation=None)
        self.density_bias = density_bias
        self.field_output_low = U.RGBFieldHead(256)           # Q18: present, never used
        self.field_output_bottleneck = U.FieldHead(out_dim=256, field_head_name="bottleneck", in_dim=256)
        self.mlp_mid = U.MLP(in_dim=34 + 256, num_layers=1, layer_width=128, out_activation=nn.ReLU())
        self.field_output_mid = U.RGBFieldHead(128)
        self.field_output_normals = U.PredNormalsFieldHead(in_dim=256, activation=None)
        self.field_output_roughness = U.FieldHead(out_dim=1, field_head_name="roughness", in_dim=256)
        self.field_output_diff = U.RGBFieldHead(256)
        self.field_output_tint = U.RGBFieldHead(256)

    # field.py:122-137
    def density(self, mean: Tensor, cov: Tensor, want_grad: bool):
        if want_grad and self.training:
            mean.requires_grad = True
            self._sample_locations = mean
        emb = self.mlp_base(ipe(mean, cov))
        raw = self.field_output_density(emb)
        if want_grad and self.training:
            self._density_before_activation = raw
        return F.softplus(raw + self.density_bias), emb

    def rgb_from(self, ide_feat: Tensor, emb: Tensor) -> Tensor:
        """field.py:167-174 tail: bottleneck -> [IDE34, b256] -> mid MLP -> sigmoid rgb."""
        bott = self.field_output_bottleneck(emb)
        return self.field_output_mid(self.mlp_mid(torch.cat([ide_feat, bott], dim=-1)))

    def point_heads(self, mean: Tensor, cov: Tensor, view_dirs: Tensor, primary: bool) -> Dict[str, Tensor]:
        """Everything the model asks of the field for one pass over samples (model.py:153-175 for the
        primary passes, model.py:295-310 for the reflected ones)."""
        sigma, emb = self.density(mean, cov, want_grad=primary)
        out = {"density": sigma, "embedding": emb}
        # field.py:139-144 (Q7): normalize(-normalize(Linear(emb)))
        out["pred_normals"] = F.normalize(-self.field_output_normals(emb), dim=-1)
        if primary:
            out["normals"] = self.get_normals() if self.training else out["pred_normals"]   # Q8
            out["n_dot_d"] = torch.sum(view_dirs * out["pred_normals"], dim=-1, keepdim=True)  # field.py:204
        out["diff"] = self.field_output_diff(emb)
        out["tint"] = self.field_output_tint(emb)
        rough_raw = self.field_output_roughness(emb)
        out["roughness_softplus"] = F.softplus(rough_raw)
        out["roughness_sigmoid"] = torch.sigmoid(rough_raw)      # Q19: second evaluation, model.py:225
        feat = ide(view_dirs, out["roughness_softplus"].detach())
        out["mid"] = self.rgb_from(feat, emb)
        out["rgb"] = out["diff"] + out["tint"] * out["mid"]
        return out

    def inf_color(self, directions: Tensor, sqradius: Tensor) -> Tensor:
        """field.py:190-201: UNcontracted gaussian at 2w, zero IDE."""
        outer = directions[..., :, None] * directions[..., None, :]
        eye = torch.eye(3, device=directions.device).expand(outer.shape)
        _, emb = self.density(2 * directions, 0.6 * sqradius[..., None] * (eye - outer), want_grad=False)
        zeros = torch.zeros(emb.shape[:-1] + (34,), device=emb.device)
        return self.rgb_from(zeros, emb)


# ----------------------------------------------------------------------------- K8
def composite(density: Tensor, starts: Tensor, ends: Tensor) -> Dict[str, Tensor]:
    """App. A.5 + A.6 for one ray batch: weights, accumulation, median depth.
    density/starts/ends: [N,S,1]."""
    rs = U.RaySamples(frustums=U.Frustums(None, None, starts, ends, None), deltas=ends - starts)
    w = rs.get_weights(density)
    return {"weights": w, "accumulation": U.AccumulationRenderer.forward(w),
            "depth": U.DepthRenderer()(w, rs)}


def blend(channels: Tensor, weights: Tensor, background, training: bool) -> Tensor:
    r = U.RGBRenderer(background_color=background)
    r.train(training)
    return r(channels, weights)


# ----------------------------------------------------------------------------- whole path
class OracleModel(nn.Module):
    """model.py:78-430 restated.  Debug prints / quantiles / .item() syncs (Q12) are dropped."""

    def __init__(self, num_coarse_samples=128, num_importance_samples=128, num_reflect_coarse_samples=64,
                 num_reflect_importance_samples=64, near_plane=2.0, far_plane=6.0) -> None:
        super().__init__()
        self.field = OracleField()
        self.sampler_uniform = make_spaced_sampler("uniform", num_coarse_samples)
        self.sampler_pdf = U.PDFSampler(num_samples=num_importance_samples, include_original=False)
        self.sampler_reciprocal = make_spaced_sampler("reciprocal", num_reflect_coarse_samples)
        self.sampler_reflect_pdf = U.PDFSampler(num_samples=num_reflect_importance_samples, include_original=False)
        self.collider = U.NearFarCollider(near_plane=near_plane, far_plane=far_plane)
        self.loss_coefficients = dict(LOSS_COEFFICIENTS)

    def set_jitter(self, uniform=None, pdf=None, reciprocal=None, reflect_pdf=None) -> None:
        """Inject the stratification noise (RNG streams cannot be matched across devices)."""
        self.sampler_uniform.injected_rand = uniform
        self.sampler_pdf.injected_rand = pdf
        self.sampler_reciprocal.injected_rand = reciprocal
        self.sampler_reflect_pdf.injected_rand = reflect_pdf

    def _pass(self, samples: U.RaySamples, primary: bool) -> Dict[str, Tensor]:
        fr = samples.frustums
        mean, cov = frustum_gaussian(fr.origins, fr.directions, fr.starts, fr.ends, fr.pixel_area)
        mean, cov = contract(mean, cov)
        out = self.field.point_heads(mean, cov, fr.directions, primary)
        out["weights"] = samples.get_weights(out["density"])
        return out

    def forward(self, bundle: U.RayBundle) -> Dict[str, Tensor]:
        return self.get_outputs(self.collider(bundle))

    def get_outputs(self, bundle: U.RayBundle) -> Dict[str, Tensor]:
        train = self.training
        white = U.WHITE.to(bundle.origins.device)
        depth_of = U.DepthRenderer()

        # A. coarse (model.py:148-177)
        su = self.sampler_uniform(bundle)
        pc = self._pass(su, primary=True)
        rgb_c = torch.clip(blend(pc["rgb"], pc["weights"], white, train), 0.0, 1.0)
        # B. fine (model.py:182-211)
        sp = self.sampler_pdf(bundle, su, pc["weights"])
        pf = self._pass(sp, primary=True)
        w = pf["weights"]
        acc_f = torch.sum(w, dim=-2)
        depth_f = depth_of(w, sp)
        rgb_f = torch.clip(blend(pf["rgb"], w, white, train), 0.0, 1.0)

        # C. per-ray quantities for the bounce (model.py:215-229)
        diff_r = blend(pf["diff"], w, white, train).detach()
        tint_r = blend(pf["tint"], w, "random", train).detach()
        nrm_r = U.NormalsRenderer.forward(pf["pred_normals"], w).detach()
        ndd = torch.sum(nrm_r * bundle.directions, dim=-1, keepdim=True).detach()
        rough = torch.sum(w * pf["roughness_sigmoid"], dim=-2)          # not detached
        mask = torch.logical_and(acc_f > 1e-2, ndd < 0).reshape(-1)

        fallback = white.expand(rgb_f.shape) * (1.0 - acc_f)           # Q10
        out = {
            "mid_rgb_coarse": rgb_c, "mid_rgb_fine": rgb_f,
            "mid_reflect_coarse": fallback, "mid_reflect_fine": white.expand(rgb_f.shape) * (1.0 - acc_f),
            "accumulation_coarse": torch.sum(pc["weights"], dim=-2).detach(),
            "accumulation_fine": acc_f.detach(),
            "depth_coarse": depth_of(pc["weights"], su).detach(), "depth_fine": depth_f.detach(),
            "weights_coarse": pc["weights"].detach(), "weights_fine": w.detach(),
            "pred_normals_coarse": pc["pred_normals"], "pred_normals_fine": pf["pred_normals"],
            "normals_coarse": pc["normals"].detach(), "normals_fine": pf["normals"].detach(),
            "n_dot_d_coarse": pc["n_dot_d"], "n_dot_d_fine": pf["n_dot_d"],
            "diff": diff_r, "tint": tint_r, "roughness": rough, "mask": mask,
        }
        if not mask.any():                                              # Q11
            return out

        # D. bounce set-up (model.py:267-290)
        o2, w_r, sqr = reflect_setup(bundle.origins[mask], bundle.directions[mask], depth_f[mask],
                                     nrm_r[mask], ndd[mask], rough[mask])
        m = o2.shape[0]
        b2 = U.RayBundle(origins=o2, directions=w_r, pixel_area=torch.pi * sqr,
                         nears=torch.zeros(m, 1, device=o2.device) * NEAR_REFLECT,      # Q4
                         fars=torch.ones(m, 1, device=o2.device) * FAR_REFLECT)
        bg = self.field.inf_color(w_r, sqr)

        # E./F. reflected coarse + fine (model.py:292-341); reflected weights are detached
        sr = self.sampler_reciprocal(b2)
        qc = self._pass(sr, primary=False)
        wc = qc["weights"].detach()
        refl_c = blend(qc["rgb"], wc, bg, train)
        out["mid_reflect_coarse"][mask, :] = diff_r[mask, :] + tint_r[mask, :] * refl_c
        out["mid_reflect_coarse"][mask, :] = torch.clip(out["mid_reflect_coarse"][mask, :], 0.0, 1.0)

        sq = self.sampler_reflect_pdf(b2, sr, wc)
        qf = self._pass(sq, primary=False)
        wf = qf["weights"].detach()
        refl_f = blend(qf["rgb"], wf, bg, train)
        out["mid_reflect_fine"][mask, :] = diff_r[mask, :] + tint_r[mask, :] * refl_f
        out["mid_reflect_fine"][mask, :] = torch.clip(out["mid_reflect_fine"][mask, :], 0.0, 1.0)
        out["depth_reflect_fine"] = depth_of(wf, sq)
        return out

    def get_loss_dict(self, outputs: Dict[str, Tensor], batch: Dict[str, Tensor]) -> Dict[str, Tensor]:
        """model.py:346-430.  blend_background_for_loss_computation is the identity for a tensor
        background and an RGB ground truth (App. A.6)."""
        image = batch["image"][..., :3]
        mse = F.mse_loss
        wc, wf = outputs["weights_coarse"], outputs["weights_fine"]
        sq = lambda a, b: torch.sum((a - b) ** 2, dim=-1, keepdim=True)  # noqa: E731
        pos = lambda v: torch.max(torch.zeros_like(v), v) ** 2            # noqa: E731
        loss = {
            "loss_mid_coarse": mse(outputs["mid_rgb_coarse"], image),
            "loss_mid_fine": mse(outputs["mid_rgb_fine"], image),
            "loss_reflect_mid_coarse": mse(outputs["mid_reflect_coarse"], image),
            "loss_reflect_mid_fine": mse(outputs["mid_reflect_fine"], image),
            "predicted_normal_loss_coarse": torch.sum(wc * sq(outputs["normals_coarse"], outputs["pred_normals_coarse"])),
            "predicted_normal_loss_fine": torch.sum(wf * sq(outputs["normals_fine"], outputs["pred_normals_fine"])),
            "orientation_loss_coarse": torch.sum(wc * pos(outputs["n_dot_d_coarse"])),
            "orientation_loss_fine": torch.sum(wf * pos(outputs["n_dot_d_fine"])),
        }
        return U.scale_dict(loss, self.loss_coefficients)


# ----------------------------------------------------------------------------- K9
def reflect_setup(origins, directions, depth, normals, n_dot_d, roughness):
    """model.py:267-272 on already-masked rows: o' = o + depth d (detached), w_r = normalize(d - 2(n.d)n)
    (detached), sqradius = 2|n.d| rho^2 (carries grad to rho, App. D)."""
    o2 = (origins + depth * directions).detach()
    w_r = F.normalize(directions - 2 * n_dot_d * normals, dim=-1).detach()
    sqr = 2 * torch.abs(n_dot_d) * roughness**2
    return o2, w_r, sqr
