"""Oracle: ThermalNerfactoModel.get_outputs / get_loss_dict as plain functions over a state_dict.

TEST INFRASTRUCTURE ONLY.  Follows models/thermal_nerfacto.py:403-489 (get_outputs),
models/nerfacto.py:299-353 (_get_outputs), model_components/ray_samplers.py:577-618
(ProposalNetworkSampler), models/thermal_nerfacto.py:253-388 (metrics + losses) and
model_components/losses.py:57-158, 593-651.  Reference quirks are reproduced, not fixed (SURVEY.md
8a notes): SH on (dir+1)/2, batch-global expected-depth clip, `removal_thermal` weights built from
the RGB samples' deltas, distortion term counted once per output suffix on the *summed* metric.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import fields as F_
from . import render as R_
from . import sampling as S_

EPS = 1.0e-7  # model_components/losses.py:39


@dataclass
class OracleConfig:
    """Mirror of ThermalNerfactoModelConfig / NerfactoModelConfig defaults
    (models/thermal_nerfacto.py:32-64, models/nerfacto.py:52-133)."""

    density_mode: str = "separate"
    num_levels: int = 16
    base_res: int = 16
    max_res: int = 2048
    log2_hashmap_size: int = 19
    num_proposal_samples_per_ray: Tuple[int, ...] = (256, 96)
    num_nerf_samples_per_ray: int = 48
    proposal_net_args_list: List[Dict] = field(default_factory=lambda: [
        {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 128},
        {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 256},
    ])
    near_plane: float = 0.05
    far_plane: float = 1000.0
    background_color: str = "last_sample"
    average_init_density: float = 1.0
    use_average_appearance_embedding: bool = True
    camera_optimizer_mode: str = "SO3xR3"
    camera_optimizer_thermal_mode: str = "SO3xR3"
    camera_penalty_scale: float = 1.0
    camera_thermal_penalty_scale: float = 10.0
    trans_l2_penalty: float = 1e-2
    rot_l2_penalty: float = 1e-3
    density_loss_mult: float = 5e-5
    rgb_density_loss_mult: float = 0.01
    thermal_loss_mult: float = 100.0
    tv_pixel_loss_mult: float = 1e-6
    cross_channel_loss_mult: float = 1e-6
    interlevel_loss_mult: float = 1.0
    distortion_loss_mult: float = 0.002
    removal_min_density_diff: float = 0.05
    is_thermal_cameras: Tuple[int, ...] = ()

    def prop_args(self, i: int) -> Dict:
        a = self.proposal_net_args_list[min(i, len(self.proposal_net_args_list) - 1)]
        return dict(num_levels=a["num_levels"], base_res=16, max_res=a["max_res"],
                    log2_hashmap_size=a["log2_hashmap_size"], average_init_density=self.average_init_density)

    def field_args(self) -> Dict:
        return dict(num_levels=self.num_levels, base_res=self.base_res, max_res=self.max_res,
                    log2_hashmap_size=self.log2_hashmap_size, average_init_density=self.average_init_density)


# ---------------------------------------------------------------- camera optimizer (SO3xR3)
def exp_map_so3xr3(tangent: torch.Tensor) -> torch.Tensor:
    """cameras/lie_groups.py:24-59 -> [B,3,4]."""
    log_rot = tangent[:, 3:]
    nrms = (log_rot * log_rot).sum(1)
    ang = torch.clamp(nrms, 1e-4).sqrt()
    inv = 1.0 / ang
    fac1 = inv * ang.sin()
    fac2 = inv * inv * (1.0 - ang.cos())
    sk = torch.zeros((log_rot.shape[0], 3, 3), dtype=log_rot.dtype, device=log_rot.device)
    sk[:, 0, 1] = -log_rot[:, 2]
    sk[:, 0, 2] = log_rot[:, 1]
    sk[:, 1, 0] = log_rot[:, 2]
    sk[:, 1, 2] = -log_rot[:, 0]
    sk[:, 2, 0] = -log_rot[:, 1]
    sk[:, 2, 1] = log_rot[:, 0]
    sk2 = torch.bmm(sk, sk)
    ret = torch.zeros(tangent.shape[0], 3, 4, dtype=tangent.dtype, device=tangent.device)
    ret[:, :3, :3] = fac1[:, None, None] * sk + fac2[:, None, None] * sk2 + torch.eye(3, device=tangent.device)[None]
    ret[:, :3, 3] = tangent[:, :3]
    return ret


def apply_camera_optimizer(pose_adjustment: torch.Tensor, frozen: torch.Tensor, camera_indices: torch.Tensor,
                           origins: torch.Tensor, directions: torch.Tensor):
    """CameraOptimizer.forward + apply_to_raybundle, cameras/camera_optimizers.py:132-176.
    frozen[num_cams] bool marks non-trainable cameras (identity correction)."""
    idx = camera_indices.squeeze()
    corr = exp_map_so3xr3(pose_adjustment[idx, :])
    corr[frozen[idx]] = torch.eye(4, device=corr.device)[:3, :4]
    origins = origins + corr[:, :3, 3]
    directions = torch.bmm(corr[:, :3, :3], directions[..., None]).squeeze()
    return origins, directions


# ---------------------------------------------------------------- proposal sampler
def _proposal_sampler(sd, cfg: OracleConfig, prop_prefix: str, origins, directions, camera_indices, nears, fars,
                      jitters: Optional[List[torch.Tensor]], updated: bool, anneal: float):
    """ProposalNetworkSampler.generate_ray_samples, ray_samplers.py:577-618."""
    weights_list, samples_list = [], []
    n = len(cfg.num_proposal_samples_per_ray)
    samples, weights = None, None
    for lvl in range(n + 1):
        is_prop = lvl < n
        num = cfg.num_proposal_samples_per_ray[lvl] if is_prop else cfg.num_nerf_samples_per_ray
        jit = None if jitters is None else jitters[lvl]
        if lvl == 0:
            samples = S_.initial_samples(origins, directions, camera_indices, nears, fars, num, jit)
        else:
            samples = S_.pdf_resample(samples, torch.pow(weights, anneal), num, jit)
        if is_prop:
            pos = S_.sample_positions(samples)
            if updated:
                dens = F_.proposal_density(sd, f"{prop_prefix}.{lvl}", pos, **cfg.prop_args(lvl))
            else:
                with torch.no_grad():
                    dens = F_.proposal_density(sd, f"{prop_prefix}.{lvl}", pos, **cfg.prop_args(lvl))
            weights = S_.sample_weights(samples.deltas, dens)
            weights_list.append(weights)
            samples_list.append(samples)
    return samples, weights_list, samples_list


def _field_forward(sd, cfg, prefix, samples: S_.OracleSamples, training: bool):
    """Field.forward, fields/base_field.py:114-133."""
    dens, geo = F_.density_field(sd, prefix, S_.sample_positions(samples), **cfg.field_args())
    dirs = samples.directions[:, None, :].expand(-1, dens.shape[1], -1)
    cams = samples.camera_indices[:, None, :].expand(-1, dens.shape[1], -1)
    col = F_.colour_head(sd, prefix, dirs, geo, cams, training=training,
                         use_average_appearance_embedding=cfg.use_average_appearance_embedding)
    return dens, col


def _get_outputs(sd, cfg, prefix, samples, weights_list, samples_list, training) -> Dict:
    """NerfactoModel._get_outputs, models/nerfacto.py:299-353."""
    dens, col = _field_forward(sd, cfg, prefix, samples, training)
    w = S_.sample_weights(samples.deltas, dens)
    weights_list.append(w)
    samples_list.append(samples)
    out = {"rgb": R_.render_colour(col, w, cfg.background_color, training)}
    with torch.no_grad():
        out["depth"] = R_.render_depth_median(w, samples.starts, samples.ends)
    out["expected_depth"] = R_.render_depth_expected(w, samples.starts, samples.ends)
    out["accumulation"] = R_.render_accumulation(w)
    out["density"] = dens
    if training:
        out["weights_list"] = weights_list
        out["ray_samples_list"] = samples_list
    for i in range(len(cfg.num_proposal_samples_per_ray)):
        out[f"prop_depth_{i}"] = R_.render_depth_median(weights_list[i], samples_list[i].starts, samples_list[i].ends)
    out["_field_rgb"] = col
    # checker-only extras (not reference outputs): what the median-depth searchsorted saw
    out["_weights"] = [x.detach() for x in weights_list]
    out["_steps"] = [((x.starts + x.ends) / 2).detach() for x in samples_list]
    return out


def thermal_nerfacto_forward(sd: Dict[str, torch.Tensor], cfg: OracleConfig, origins: torch.Tensor,
                             directions: torch.Tensor, camera_indices: torch.Tensor, *, training: bool,
                             jitters: Optional[List[torch.Tensor]] = None,
                             jitters_thermal: Optional[List[torch.Tensor]] = None,
                             updated: bool = True, anneal: float = 1.0) -> Dict:
    """Model.forward (collider) + ThermalNerfactoModel.get_outputs.

    jitters / jitters_thermal: the [R,1] torch.rand draws of the three samplers of each path in
    call order (RGB path first, thermal second; SURVEY.md section 7 "RNG parity").  None in training
    means "draw them here with torch.rand in the reference's order"; ignored in eval.
    """
    R = origins.shape[0]
    near = cfg.near_plane if training else 0.0  # scene_colliders.py:186-191
    nears = torch.ones_like(origins[..., 0:1]) * near
    fars = torch.ones_like(origins[..., 0:1]) * cfg.far_plane
    n_lvls = len(cfg.num_proposal_samples_per_ray) + 1
    frozen_rgb = torch.tensor([bool(t) for t in cfg.is_thermal_cameras], device=origins.device)
    frozen_thermal = ~frozen_rgb

    def _draw(given):
        if not training:
            return None
        return given  # may be None -> drawn lazily below

    class _Lazy(list):
        """draws torch.rand(R,1) on first access of each level, preserving the reference's RNG order"""

        def __getitem__(self, i):
            while len(self) <= i:
                self.append(torch.rand((R, 1)).to(origins.device))
            return list.__getitem__(self, i)

    jit = _draw(jitters)
    if training and jit is None:
        jit = _Lazy()

    o, d = origins, directions
    if training and cfg.camera_optimizer_mode != "off":
        o, d = apply_camera_optimizer(sd["camera_optimizer.pose_adjustment"], frozen_rgb, camera_indices, o, d)
    samples, wl, sl = _proposal_sampler(sd, cfg, "proposal_networks", o, d, camera_indices, nears, fars, jit,
                                        updated, anneal)
    out = _get_outputs(sd, cfg, "field", samples, wl, sl, training)
    field_rgb = out.pop("_field_rgb")

    if cfg.density_mode == "shared":
        rgbt = out["rgb"]
        out["rgbt"] = rgbt
        out["rgb"] = rgbt[..., :3]
        out["rgb_thermal"] = rgbt[..., 3:]
    elif cfg.density_mode == "separate":
        jit_t = _draw(jitters_thermal)
        if training and jit_t is None:
            jit_t = _Lazy()
        ot, dt = origins, directions
        if training and cfg.camera_optimizer_thermal_mode != "off":
            ot, dt = apply_camera_optimizer(sd["camera_optimizer_thermal.pose_adjustment"], frozen_thermal,
                                            camera_indices, ot, dt)
        samples_t, wl_t, sl_t = _proposal_sampler(sd, cfg, "proposal_networks_thermal", ot, dt, camera_indices, nears,
                                                  fars, jit_t, updated, anneal)
        out_t = _get_outputs(sd, cfg, "field_thermal", samples_t, wl_t, sl_t, training)
        field_rgb_t = out_t.pop("_field_rgb")
        for k, v in out_t.items():
            out[f"{k}_thermal"] = v
        if cfg.density_loss_mult > 0 or not training:
            out["density2"], _ = _field_forward(sd, cfg, "field", samples_t, training)
            out["density2_thermal"], _ = _field_forward(sd, cfg, "field_thermal", samples, training)
        if not training:  # thermal_nerfacto.py:460-487 (the two field forwards there recompute field_rgb*)
            thr = cfg.removal_min_density_diff
            m = (out["density"] / out["density"] - out["density2_thermal"] / out["density"]).abs() < thr
            w_rm = S_.sample_weights(samples.deltas, out["density"] * m)
            out["removal"] = R_.render_colour(field_rgb, w_rm, cfg.background_color, training)
            m_t = (out["density_thermal"] / out["density_thermal"] - out["density2"] / out["density_thermal"]).abs() < thr
            w_rm_t = S_.sample_weights(samples.deltas, out["density_thermal"] * m_t)  # RGB deltas: reference quirk
            out["removal_thermal"] = R_.render_colour(field_rgb_t, w_rm_t, cfg.background_color, training)
    return out


# ---------------------------------------------------------------- losses
def _outer(t0_starts, t0_ends, t1_starts, t1_ends, y1):
    """losses.py:57-84."""
    cy1 = torch.cat([torch.zeros_like(y1[..., :1]), torch.cumsum(y1, dim=-1)], dim=-1)
    lo = torch.searchsorted(t1_starts.contiguous(), t0_starts.contiguous(), side="right") - 1
    lo = torch.clamp(lo, min=0, max=y1.shape[-1] - 1)
    hi = torch.searchsorted(t1_ends.contiguous(), t0_ends.contiguous(), side="right")
    hi = torch.clamp(hi, min=0, max=y1.shape[-1] - 1)
    return torch.take_along_dim(cy1[..., 1:], hi, dim=-1) - torch.take_along_dim(cy1[..., :-1], lo, dim=-1)


def interlevel_loss(weights_list, samples_list) -> torch.Tensor:
    """losses.py:87-135."""
    c = samples_list[-1].sdist().detach()
    w = weights_list[-1][..., 0].detach()
    total = 0.0
    for s, wp in zip(samples_list[:-1], weights_list[:-1]):
        cp = s.sdist()
        w_outer = _outer(c[..., :-1], c[..., 1:], cp[..., :-1], cp[..., 1:], wp[..., 0])
        total = total + torch.mean(torch.clip(w - w_outer, min=0) ** 2 / (w + EPS))
    return total


def distortion_loss(weights_list, samples_list) -> torch.Tensor:
    """losses.py:139-158."""
    t = samples_list[-1].sdist()
    w = weights_list[-1][..., 0]
    ut = (t[..., 1:] + t[..., :-1]) / 2
    dut = torch.abs(ut[..., :, None] - ut[..., None, :])
    inter = torch.sum(w * torch.sum(w[..., None, :] * dut, dim=-1), dim=-1)
    intra = torch.sum(w**2 * (t[..., 1:] - t[..., :-1]), dim=-1) / 3
    return torch.mean(inter + intra)


def _rgb_to_rgbt(image, is_thermal):
    """utils/rgbt_utils.py:6-33."""
    rgbt = torch.zeros(image.shape[:-1] + (4,), device=image.device)
    rgbt[..., :3] = torch.einsum("ij,i->ij", image, 1 - is_thermal)
    rgbt[..., 3] = image[..., 0] * is_thermal
    return rgbt


def tv_pixel_loss(pred_thermal, is_thermal):
    """losses.py:603-620."""
    p = pred_thermal[(1 - is_thermal).bool()].view(-1, 4)
    return 0.25 * torch.mean((p[:, 0] - p[:, 1]).abs() + (p[:, 0] - p[:, 2]).abs()
                             + (p[:, 1] - p[:, 3]).abs() + (p[:, 2] - p[:, 3]).abs())


def _pixel_grad(img):
    p = img.view(-1, 4)
    return torch.stack((p[:, 1] - p[:, 0], p[:, 2] - p[:, 0], p[:, 3] - p[:, 1], p[:, 3] - p[:, 2]))


def cross_channel_loss(pred_thermal, gt_rgb, is_thermal):
    """losses.py:623-651."""
    keep = (1 - is_thermal).bool()
    diff = (_pixel_grad(pred_thermal[keep]) - _pixel_grad(gt_rgb[keep].mean(-1, keepdim=True))).abs()
    return 0.25 * (diff[0, :] + diff[1, :] + diff[2, :] + diff[3, :]).mean()


def _camera_reg(pose, cfg, scale):
    """camera_optimizers.py:189-195."""
    return (pose[:, :3].norm(dim=-1).mean() * cfg.trans_l2_penalty
            + pose[:, 3:].norm(dim=-1).mean() * cfg.rot_l2_penalty) * scale


def thermal_nerfacto_losses(sd, cfg: OracleConfig, outputs: Dict, image: torch.Tensor, is_thermal: torch.Tensor,
                            *, training: bool = True) -> Dict[str, torch.Tensor]:
    """get_metrics_dict["distortion"] + get_loss_dict, models/thermal_nerfacto.py:253-388."""
    mse = torch.nn.functional.mse_loss
    l1 = torch.nn.functional.l1_loss
    loss = {}
    suffixes = ("", "_thermal") if cfg.density_mode == "separate" else ("",)
    if cfg.density_mode != "rgb_only":
        pred = torch.cat((outputs["rgb"], outputs["rgb_thermal"]), dim=1)
    else:
        pred = torch.cat((outputs["rgb"], torch.zeros(outputs["rgb"].shape[0], 1, device=outputs["rgb"].device)), dim=1)
    if cfg.background_color == "random":
        raise NotImplementedError("oracle covers last_sample/black/white backgrounds")
    gt = _rgb_to_rgbt(image, is_thermal)
    is_rgb = (1 - is_thermal)[:, None]
    loss["rgb_loss"] = mse(gt[..., :3] * is_rgb, pred[..., :3] * is_rgb)
    if cfg.density_mode != "rgb_only":
        th = is_thermal[:, None]
        loss["thermal_loss"] = cfg.thermal_loss_mult * mse(gt[..., 3:] * th, pred[..., 3:] * th)
    if cfg.density_mode == "separate" and cfg.density_loss_mult > 0:
        m, r = cfg.density_loss_mult, cfg.rgb_density_loss_mult
        if r == 1:
            dl = m * l1(outputs["density2"], outputs["density_thermal"])
            dl = dl + m * l1(outputs["density"], outputs["density2_thermal"])
        else:
            dl = m * l1(outputs["density2"].detach(), outputs["density_thermal"])
            dl = dl + m * l1(outputs["density"].detach(), outputs["density2_thermal"])
            dl = dl + r * m * l1(outputs["density2"], outputs["density_thermal"].detach())
            dl = dl + r * m * l1(outputs["density"], outputs["density2_thermal"].detach())
        loss["density_loss"] = dl
    if cfg.density_mode != "rgb_only" and cfg.tv_pixel_loss_mult > 0:
        loss["tv_pixel_loss"] = cfg.tv_pixel_loss_mult * tv_pixel_loss(pred[..., 3:], is_thermal)
    if cfg.density_mode != "rgb_only" and cfg.cross_channel_loss_mult > 0:
        loss["cross_channel_loss"] = cfg.cross_channel_loss_mult * cross_channel_loss(pred[..., 3:], gt[..., :3],
                                                                                      is_thermal)
    if training:
        distortion = 0
        for s in suffixes:
            distortion = distortion + distortion_loss(outputs[f"weights_list{s}"], outputs[f"ray_samples_list{s}"])
        loss["interlevel_loss"] = 0
        loss["distortion_loss"] = 0
        for s in suffixes:
            loss["interlevel_loss"] = loss["interlevel_loss"] + cfg.interlevel_loss_mult * interlevel_loss(
                outputs[f"weights_list{s}"], outputs[f"ray_samples_list{s}"])
            loss["distortion_loss"] = loss["distortion_loss"] + cfg.distortion_loss_mult * distortion
        if cfg.camera_optimizer_mode != "off":
            loss["camera_opt_regularizer"] = _camera_reg(sd["camera_optimizer.pose_adjustment"], cfg,
                                                         cfg.camera_penalty_scale)
        if cfg.density_mode == "separate" and cfg.camera_optimizer_thermal_mode != "off":
            loss["camera_opt_regularizer_thermal"] = _camera_reg(sd["camera_optimizer_thermal.pose_adjustment"], cfg,
                                                                 cfg.camera_thermal_penalty_scale)
    return loss
