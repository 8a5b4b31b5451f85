"""ThermalNerfactoModel: the caller of the hot path, with the reference's structure, output keys,
loss terms, parameter groups and state_dict keys.

Mirrors `nerfstudio/models/thermal_nerfacto.py:32-489`, `nerfstudio/models/nerfacto.py:52-353`,
`nerfstudio/models/base_model.py:132-206`, `nerfstudio/model_components/scene_colliders.py:169-191` and
`nerfstudio/cameras/camera_optimizers.py:89-213` (SO3xR3 mode).  Data managers, trainers, viewers and
metrics (PSNR/SSIM/LPIPS) are the reference's host code and are out of scope.
"""
import os
from collections import defaultdict
from dataclasses import dataclass, field
from typing import Dict, List, Literal, Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn
from torch.nn import Parameter

from . import fused_ops, ops
from .field_components import SceneContraction
from .fields import FieldHeadNames, HashMLPDensityField, ThermalNerfactoField
from .losses import L1Loss, MSELoss, cross_channel_loss, distortion_loss, interlevel_loss, tv_pixel_loss
from .rays import RayBundle, RayLayout, RaySamples, samples_from_layout
from .renderers import AccumulationRenderer, DepthRenderer, RGBRenderer, RGBTRenderer
from .samplers import (PDFSampler, ProposalNetworkSampler, UniformLinDispPiecewiseSampler, UniformSampler,
                       _PiecewiseSpacing)


# ------------------------------------------------------------------------------------------ camera optimizer
@dataclass
class CameraOptimizerConfig:
    """cameras/camera_optimizers.py:39-84 (fields the model reads)."""

    mode: Literal["off", "SO3xR3", "shared_SO3xR3"] = "off"
    trans_l2_penalty: float = 1e-2
    rot_l2_penalty: float = 1e-3
    penalty_scale: float = 1


def exp_map_SO3xR3(tangent_vector: Tensor) -> Tensor:
    """[B,6] (translation, so(3) log-rotation) -> [B,3,4] = [R|t].  cameras/lie_groups.py:24-59 (Rodrigues)."""
    t, w = tangent_vector[:, :3], tangent_vector[:, 3:]
    theta = torch.clamp((w * w).sum(1), 1e-4).sqrt()
    inv = 1.0 / theta
    a = inv * theta.sin()
    b = inv * inv * (1.0 - theta.cos())
    zero = torch.zeros_like(w[:, 0])
    k = torch.stack([zero, -w[:, 2], w[:, 1], w[:, 2], zero, -w[:, 0], -w[:, 1], w[:, 0], zero], dim=1).view(-1, 3, 3)
    rot = a[:, None, None] * k + b[:, None, None] * torch.bmm(k, k) + torch.eye(3, dtype=w.dtype, device=w.device)[None]
    return torch.cat([rot, t[:, :, None]], dim=2)


class CameraOptimizer(nn.Module):
    """cameras/camera_optimizers.py:89-213."""

    def __init__(self, config: CameraOptimizerConfig, num_cameras: int, device="cpu",
                 non_trainable_camera_indices: Optional[Tensor] = None, suffix: str = "") -> None:
        super().__init__()
        self.config = CameraOptimizerConfig(**vars(config))
        self.num_cameras = num_cameras
        if self.config.penalty_scale < 0:
            self.config.mode = "off"
        if self.config.mode == "SO3xR3":
            self.pose_adjustment = Parameter(torch.zeros((num_cameras, 6), device=device))
        elif self.config.mode == "shared_SO3xR3":
            self.pose_adjustment = Parameter(torch.zeros((1, 6), device=device))
        elif self.config.mode != "off":
            raise NotImplementedError(f"camera optimizer mode {self.config.mode}")
        self.suffix = suffix
        frozen = torch.zeros(num_cameras, dtype=torch.bool)
        if non_trainable_camera_indices is not None and len(non_trainable_camera_indices) > 0:
            frozen[non_trainable_camera_indices.long()] = True
        self.has_frozen = non_trainable_camera_indices is not None
        self.register_buffer("_frozen", frozen, persistent=False)
        self.register_buffer("_frozen_u8", frozen.to(torch.uint8), persistent=False)
        self.fused = True  # single-kernel apply_to_raybundle on CUDA (the torch expression is kept in forward())
        self._reg_cache = None
        # accumulation target of d(pose_adjustment) for the fused kernels (parallel.FlatGradBuffer.attach_sinks)
        self.grad_sink: Optional[Tensor] = None

    def forward(self, indices: Tensor) -> Tensor:
        if self.config.mode == "off":
            return torch.eye(4, device=indices.device)[None, :3, :4].tile(indices.shape[0], 1, 1)
        if self.config.mode == "SO3xR3":
            out = exp_map_SO3xR3(self.pose_adjustment[indices, :])
        else:
            out = exp_map_SO3xR3(self.pose_adjustment).tile((indices.shape[0], 1, 1))
        if self.has_frozen:  # non-trainable cameras get the identity (camera_optimizers.py:155-164)
            eye = torch.eye(4, device=out.device)[:3, :4]
            out = torch.where(self._frozen[indices][:, None, None], eye, out)
        return out

    def apply_to_raybundle(self, raybundle: RayBundle) -> None:
        if self.config.mode != "off" and self.fused and raybundle.origins.is_cuda and raybundle.origins.dim() == 2:
            # exp map + masking + translation + rotation of the bundle as one kernel (and one for the backward)
            raybundle.origins, raybundle.directions = fused_ops.camera_opt_apply(
                self.pose_adjustment, self._frozen_u8 if self.has_frozen else None,
                raybundle.camera_indices.reshape(-1), raybundle.origins, raybundle.directions,
                self.config.mode == "shared_SO3xR3", sink=self.grad_sink)
            return
        if self.config.mode != "off":
            c = self(raybundle.camera_indices.squeeze(-1))
            raybundle.origins = raybundle.origins + c[:, :3, 3]
            raybundle.directions = torch.bmm(c[:, :3, :3], raybundle.directions[..., None]).squeeze(-1)

    def _reg_and_norms(self):
        """(regulariser, |t|_F, |w|_F) of the pose table from one launch (fused_ops.camera_regularizer)."""
        c = self.config
        return fused_ops.camera_regularizer(self.pose_adjustment, c.trans_l2_penalty, c.rot_l2_penalty, c.penalty_scale,
                                            sink=self.grad_sink)

    def get_loss_dict(self, loss_dict: dict) -> None:
        if self.config.mode != "off" and self.fused and self.pose_adjustment.is_cuda:
            # get_metrics_dict of the same step (get_train_loss_dict calls it first) already launched the kernel
            cached, self._reg_cache = self._reg_cache, None
            reg = cached[0] if (cached is not None and cached[0].requires_grad == torch.is_grad_enabled()) \
                else self._reg_and_norms()[0]
            loss_dict[f"camera_opt_regularizer{self.suffix}"] = reg
            return
        if self.config.mode != "off":
            loss_dict[f"camera_opt_regularizer{self.suffix}"] = (
                self.pose_adjustment[:, :3].norm(dim=-1).mean() * self.config.trans_l2_penalty
                + self.pose_adjustment[:, 3:].norm(dim=-1).mean() * self.config.rot_l2_penalty
            ) * self.config.penalty_scale

    def get_metrics_dict(self, metrics_dict: dict) -> None:
        if self.config.mode != "off" and self.fused and self.pose_adjustment.is_cuda:
            self._reg_cache = self._reg_and_norms()
            _, tn_, rn_ = self._reg_cache
            metrics_dict[f"camera_opt_translation{self.suffix}"] = tn_
            metrics_dict[f"camera_opt_rotation{self.suffix}"] = rn_
            return
        if self.config.mode != "off":
            metrics_dict[f"camera_opt_translation{self.suffix}"] = self.pose_adjustment[:, :3].norm()
            metrics_dict[f"camera_opt_rotation{self.suffix}"] = self.pose_adjustment[:, 3:].norm()

    def get_param_groups(self, param_groups: dict, name: str = "camera_opt") -> None:
        params = list(self.parameters())
        if self.config.mode != "off":
            assert len(params) > 0
            param_groups[name] = params


class NearFarCollider(nn.Module):
    """model_components/scene_colliders.py:169-191."""

    def __init__(self, near_plane: float, far_plane: float, reset_near_plane: bool = True) -> None:
        super().__init__()
        self.near_plane, self.far_plane, self.reset_near_plane = near_plane, far_plane, reset_near_plane
        self._planes: Dict = {}

    def forward(self, ray_bundle: RayBundle) -> RayBundle:
        near_plane = self.near_plane if (self.training or not self.reset_near_plane) else 0
        o = ray_bundle.origins
        # constants: built once per (batch shape, device, planes) instead of ones_like + two multiplies per call.
        # Entries are never evicted -- a captured CUDA graph reads them by ADDRESS, it holds no reference -- so only
        # training-sized batches are kept (<= 65536 rays = 0.5 MB per entry); frames and sweeps build theirs per call.
        key = (tuple(o.shape[:-1]), str(o.device), o.dtype, float(near_plane), float(self.far_plane))
        planes = self._planes.get(key)
        if planes is None:
            planes = (torch.full((*o.shape[:-1], 1), float(near_plane), device=o.device, dtype=o.dtype),
                      torch.full((*o.shape[:-1], 1), float(self.far_plane), device=o.device, dtype=o.dtype))
            # (a tensor first made while a CUDA graph is being captured only holds its values after a replay, and an
            # inference-mode tensor could not be used by a later training forward under autograd)
            if planes[0].numel() <= 65536 and not (o.is_cuda and torch.cuda.is_current_stream_capturing()) \
                    and not torch.is_inference_mode_enabled():
                self._planes[key] = planes
        ray_bundle.nears, ray_bundle.fars = planes
        return ray_bundle


# ------------------------------------------------------------------------------------------ config
@dataclass
class ThermalNerfactoModelConfig:
    """models/thermal_nerfacto.py:32-64 on top of models/nerfacto.py:52-133 (same names and defaults)."""

    near_plane: float = 0.05
    far_plane: float = 1000.0
    background_color: Literal["random", "last_sample", "black", "white"] = "last_sample"
    hidden_dim: int = 64
    hidden_dim_color: int = 64
    hidden_dim_transient: int = 64
    num_levels: int = 16
    base_res: int = 16
    max_res: int = 2048
    log2_hashmap_size: int = 19
    features_per_level: int = 2
    num_proposal_samples_per_ray: Tuple[int, ...] = (256, 96)
    num_nerf_samples_per_ray: int = 48
    proposal_update_every: int = 5
    proposal_warmup: int = 5000
    num_proposal_iterations: int = 2
    use_same_proposal_network: bool = False
    proposal_net_args_list: List[Dict] = field(default_factory=lambda: [
        {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 128, "use_linear": False},
        {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 256, "use_linear": False},
    ])
    proposal_initial_sampler: Literal["piecewise", "uniform"] = "piecewise"
    interlevel_loss_mult: float = 1.0
    distortion_loss_mult: float = 0.002
    use_proposal_weight_anneal: bool = True
    use_appearance_embedding: bool = True
    use_average_appearance_embedding: bool = True
    proposal_weights_anneal_slope: float = 10.0
    proposal_weights_anneal_max_num_iters: int = 1000
    use_single_jitter: bool = True
    predict_normals: bool = False
    disable_scene_contraction: bool = False
    use_gradient_scaling: bool = False
    implementation: Literal["tcnn", "torch", "b200"] = "b200"
    appearance_embed_dim: int = 32
    average_init_density: float = 1.0
    camera_optimizer: CameraOptimizerConfig = field(default_factory=lambda: CameraOptimizerConfig(mode="SO3xR3"))
    eval_num_rays_per_chunk: int = 1 << 15  # configs/method_configs.py:270
    # thermal
    density_loss_mult: float = 5e-5
    density_mode: Literal["rgb_only", "shared", "separate"] = "separate"
    rgb_density_loss_mult: float = 0.01
    thermal_loss_mult: float = 100.0
    tv_rgb_loss_mult: float = 0
    tv_thermal_loss_mult: float = 0
    tv_pixel_loss_mult: float = 1e-6
    cross_channel_loss_mult: float = 1e-6
    removal_min_density_diff: float = 0.05
    use_proposal_thermal_weight_anneal: bool = False
    camera_optimizer_thermal: CameraOptimizerConfig = field(
        default_factory=lambda: CameraOptimizerConfig(mode="SO3xR3", penalty_scale=10))
    shared_camera_optimizer: CameraOptimizerConfig = field(
        default_factory=lambda: CameraOptimizerConfig(mode="shared_SO3xR3", penalty_scale=-1))
    shared_camera_optimizer_thermal: CameraOptimizerConfig = field(
        default_factory=lambda: CameraOptimizerConfig(mode="shared_SO3xR3", penalty_scale=-1))

    def setup(self, **kwargs) -> "ThermalNerfactoModel":
        return ThermalNerfactoModel(self, **kwargs)


class _BranchSamples:
    """One branch between its proposal phase and its render phase (ThermalNerfactoModel.sample_phase)."""

    def __init__(self, ray_bundle: RayBundle, sampler) -> None:
        self.ray_bundle = ray_bundle  # after the pose corrections: origins / directions carry the gradient to the poses
        self.sampler = sampler
        self.generic = None           # (ray_samples, weights_list, ray_samples_list) of the component-by-component path
        self.spacing = None
        self.counts: List[int] = []
        self.ebins = self.sbins = self.nears = self.fars = None  # final level
        self.weights_list: List[Tensor] = []
        self.ray_samples_list: List[RaySamples] = []
        self.prop_sigma: List[Tensor] = []   # proposal densities [R,S_i]: the interlevel loss reaches the networks here
        self.prop_depths: List[Tensor] = []
        self.grad_props = False

    def cut(self) -> List[Tuple[Tensor, Tensor]]:
        """Replace every differentiable tensor the render phase reads by a detached leaf and return the
        (original, leaf) pairs: a backward of the render phase then stops at the leaves, and
        `torch.autograd.backward(originals, [leaf.grad ...])` finishes the job later (engine.GraphedTrainStep runs
        the gradient exchange of the main fields in between)."""
        if self.generic is not None:
            raise NotImplementedError("the two-phase backward needs the per-level launches (fuse_levels)")
        pairs: List[Tuple[Tensor, Tensor]] = []

        def leaf(t: Tensor) -> Tensor:
            if not t.requires_grad:
                return t
            d = t.detach().requires_grad_(True)
            pairs.append((t, d))
            return d

        rb = self.ray_bundle
        self.ray_bundle = RayBundle(origins=leaf(rb.origins), directions=leaf(rb.directions), pixel_area=rb.pixel_area,
                                    camera_indices=rb.camera_indices, nears=rb.nears, fars=rb.fars,
                                    metadata=rb.metadata, times=rb.times)
        self.prop_sigma = [leaf(t) for t in self.prop_sigma]
        return pairs


class _ForwardState:
    """sample_phase -> render_phase hand-over."""

    def __init__(self) -> None:
        self.rgb: Optional[_BranchSamples] = None
        self.thermal: Optional[_BranchSamples] = None
        self.streams = False

    def cut(self) -> List[Tuple[Tensor, Tensor]]:
        return [p for b in (self.rgb, self.thermal) if b is not None for p in b.cut()]


class LossDict(dict):
    """The loss dictionary of get_loss_dict plus `.total`, the sum of its values (engine/trainer.py:479)."""

    total: Optional[Tensor] = None


def _assemble_losses(entries, fused: bool) -> LossDict:
    tensors = [t for _, t, _ in entries if torch.is_tensor(t)]
    if fused and len(tensors) == len(entries) and all(t.is_cuda and t.numel() == 1 for t in tensors) \
            and 1 <= len(entries) <= 16:
        total, values = fused_ops.loss_sum([(k, t.reshape(()), w) for k, t, w in entries])
        out = LossDict(values)
        out.total = total
        return out
    out = LossDict()
    for k, t, w in entries:
        out[k] = out.get(k, 0) + w * t
    out.total = sum(out.values())
    return out


# ------------------------------------------------------------------------------------------ model
class ThermalNerfactoModel(nn.Module):
    """models/thermal_nerfacto.py:67-489 (+ the NerfactoModel / Model parts it inherits)."""

    def __init__(self, config: ThermalNerfactoModelConfig, scene_box=None, num_train_data: int = 0,
                 metadata: Optional[Dict] = None, aabb: Optional[Tensor] = None, **kwargs) -> None:
        super().__init__()
        if config.predict_normals or config.use_gradient_scaling or config.tv_rgb_loss_mult > 0 \
                or config.tv_thermal_loss_mult > 0:
            raise NotImplementedError("predict_normals / gradient scaling / density-TV are off the default path")
        self.config = config
        if aabb is None:
            aabb = scene_box.aabb if scene_box is not None else torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
        self.scene_box = scene_box
        self.num_train_data = num_train_data
        self.kwargs = dict(metadata=metadata or {}, **kwargs)
        self.device_indicator_param = nn.Parameter(torch.empty(0))  # models/base_model.py:85
        self.fuse_losses = True  # pixel / density loss terms as single kernels (torch expressions otherwise)
        # one launch per sampling level (csrc/tn_level.cu): weights + resampling per proposal level, weights + all
        # renderers + distortion + interlevel loss for the final level, one launch for the ray-level backward
        self.fuse_levels = os.environ.get("TN_FUSE_LEVELS", "1") == "1"
        self.branch_streams = os.environ.get("TN_BRANCH_STREAMS", "1") == "1"
        # eval: off for eager chunk loops (GPU-bound per kernel, measured 4 % slower), on inside the captured
        # chunk graph of engine.GraphedRenderChunk
        self.eval_branch_streams = False
        self._side_stream = None
        self._scratch = fused_ops.LaunchScratch()  # self-re-arming launch words of the per-level / loss kernels
        self._populate(aabb)

    @property
    def device(self):
        return self.device_indicator_param.device

    def _make_field(self, aabb, contraction, num_channels):
        c = self.config
        return ThermalNerfactoField(
            aabb, hidden_dim=c.hidden_dim, num_levels=c.num_levels, max_res=c.max_res, base_res=c.base_res,
            features_per_level=c.features_per_level, log2_hashmap_size=c.log2_hashmap_size,
            hidden_dim_color=c.hidden_dim_color, hidden_dim_transient=c.hidden_dim_transient,
            spatial_distortion=contraction, num_images=self.num_train_data, use_pred_normals=c.predict_normals,
            use_average_appearance_embedding=c.use_average_appearance_embedding,
            appearance_embedding_dim=c.appearance_embed_dim if c.use_appearance_embedding else 0,
            implementation=c.implementation, num_channels=num_channels)

    def _make_proposals(self, aabb, contraction):
        c = self.config
        nets, fns = torch.nn.ModuleList(), []
        if c.use_same_proposal_network:
            assert len(c.proposal_net_args_list) == 1, "Only one proposal network is allowed."
            net = HashMLPDensityField(aabb, spatial_distortion=contraction, **c.proposal_net_args_list[0],
                                      average_init_density=c.average_init_density, implementation=c.implementation)
            nets.append(net)
            fns.extend([net.density_fn for _ in range(c.num_proposal_iterations)])
        else:
            for i in range(c.num_proposal_iterations):
                args = c.proposal_net_args_list[min(i, len(c.proposal_net_args_list) - 1)]
                nets.append(HashMLPDensityField(aabb, spatial_distortion=contraction, **args,
                                                average_init_density=c.average_init_density,
                                                implementation=c.implementation))
            fns.extend([n.density_fn for n in nets])
        return nets, fns

    def _make_sampler(self):
        c = self.config

        def update_schedule(step):
            return np.clip(np.interp(step, [0, c.proposal_warmup], [0, c.proposal_update_every]), 1,
                           c.proposal_update_every)

        initial = UniformSampler(single_jitter=c.use_single_jitter) if c.proposal_initial_sampler == "uniform" else None
        return ProposalNetworkSampler(
            num_nerf_samples_per_ray=c.num_nerf_samples_per_ray,
            num_proposal_samples_per_ray=c.num_proposal_samples_per_ray,
            num_proposal_network_iterations=c.num_proposal_iterations, single_jitter=c.use_single_jitter,
            update_sched=update_schedule, initial_sampler=initial)

    def _populate(self, aabb: Tensor) -> None:
        """populate_modules: models/nerfacto.py:145-254 then models/thermal_nerfacto.py:79-216 (the modules
        the thermal subclass overwrites -- field, camera_optimizer -- are built once)."""
        c = self.config
        contraction = None if c.disable_scene_contraction else SceneContraction(order=float("inf"))
        is_thermal = list(self.kwargs["metadata"].get("is_thermal", [0] * self.num_train_data))
        thermal_idx = torch.tensor([i for i, x in enumerate(is_thermal) if x != 0], dtype=torch.long)
        rgb_idx = torch.tensor([i for i, x in enumerate(is_thermal) if x == 0], dtype=torch.long)
        # --- NerfactoModel.populate_modules
        self.proposal_networks, self.density_fns = self._make_proposals(aabb, contraction)
        self.proposal_sampler = self._make_sampler()
        self.collider = NearFarCollider(near_plane=c.near_plane, far_plane=c.far_plane)
        self.renderer_rgb = RGBRenderer(background_color=c.background_color)
        self.renderer_accumulation = AccumulationRenderer()
        self.renderer_depth = DepthRenderer(method="median")
        self.renderer_expected_depth = DepthRenderer(method="expected")
        self.rgb_loss = MSELoss()
        self.step = 0
        # --- ThermalNerfactoModel.populate_modules
        self.output_suffixes = ("", "_thermal") if c.density_mode == "separate" else ("",)
        self.field = self._make_field(aabb, contraction, 3 + (c.density_mode == "shared"))
        if c.density_mode == "separate":
            self.field_thermal = self._make_field(aabb, contraction, 1)
        self.camera_optimizer = CameraOptimizer(c.camera_optimizer, self.num_train_data,
                                                non_trainable_camera_indices=thermal_idx)
        self.camera_optimizer_thermal = CameraOptimizer(c.camera_optimizer_thermal, self.num_train_data,
                                                        non_trainable_camera_indices=rgb_idx, suffix="_thermal")
        self.shared_camera_optimizer = CameraOptimizer(c.shared_camera_optimizer, self.num_train_data,
                                                       non_trainable_camera_indices=thermal_idx, suffix="_shared")
        self.shared_camera_optimizer_thermal = CameraOptimizer(
            c.shared_camera_optimizer_thermal, self.num_train_data, non_trainable_camera_indices=rgb_idx,
            suffix="_shared_thermal")
        self.proposal_networks_thermal, self.density_fns_thermal = self._make_proposals(aabb, contraction)
        self.proposal_sampler_thermal = self._make_sampler()
        self.renderer_rgbt = RGBTRenderer(background_color=c.background_color)
        self.renderer_thermal = RGBRenderer(background_color=c.background_color, num_channels=1)
        self.density_loss = L1Loss()

    # -------------------------------------------------------------------------------------- optimisation glue
    def get_param_groups(self) -> Dict[str, List[Parameter]]:
        """models/nerfacto.py:256-261 + models/thermal_nerfacto.py:390-401."""
        groups: Dict[str, List[Parameter]] = {}
        groups["proposal_networks"] = list(self.proposal_networks.parameters())
        groups["fields"] = list(self.field.parameters())
        self.camera_optimizer.get_param_groups(param_groups=groups)
        self.shared_camera_optimizer.get_param_groups(param_groups=groups, name="shared_camera_opt")
        if self.config.density_mode == "separate":
            groups["proposal_networks_thermal"] = list(self.proposal_networks_thermal.parameters())
            groups["fields_thermal"] = list(self.field_thermal.parameters())
            self.camera_optimizer_thermal.get_param_groups(param_groups=groups, name="camera_opt_thermal")
            self.shared_camera_optimizer_thermal.get_param_groups(param_groups=groups,
                                                                  name="shared_camera_opt_thermal")
        return groups

    def set_anneal_step(self, step: int) -> None:
        """The BEFORE_TRAIN_ITERATION callback body, models/nerfacto.py:271-281."""
        c = self.config
        self.step = step
        if c.use_proposal_weight_anneal:
            frac = np.clip(step / c.proposal_weights_anneal_max_num_iters, 0, 1)
            b = c.proposal_weights_anneal_slope
            anneal = b * frac / ((b - 1) * frac + 1)
            self.proposal_sampler.set_anneal(anneal)
            if c.use_proposal_thermal_weight_anneal:
                self.proposal_sampler_thermal.set_anneal(anneal)

    def step_cb(self, step: int) -> None:
        """The AFTER_TRAIN_ITERATION callbacks (models/nerfacto.py:289-295, models/thermal_nerfacto.py:244-250): the
        proposal samplers count the iterations since their networks were last updated.  As in the reference, the
        thermal sampler is only driven when `use_proposal_thermal_weight_anneal` is on (otherwise its `_step`
        stays 0 and its proposal networks are updated on every iteration)."""
        c = self.config
        if c.use_proposal_weight_anneal:
            self.proposal_sampler.step_cb(step)
            if c.use_proposal_thermal_weight_anneal:
                self.proposal_sampler_thermal.step_cb(step)

    # -------------------------------------------------------------------------------------- forward
    def forward(self, ray_bundle: RayBundle, jitters: Optional[List[Tensor]] = None,
                jitters_thermal: Optional[List[Tensor]] = None) -> Dict[str, Tensor]:
        """Model.forward, models/base_model.py:132-143: collider, then get_outputs."""
        ray_bundle = self.collider(ray_bundle)
        return self.get_outputs(ray_bundle, jitters=jitters, jitters_thermal=jitters_thermal)

    def _get_outputs(self, ray_bundle: RayBundle, field_: ThermalNerfactoField, renderer: nn.Module,
                     ray_samples: RaySamples, weights_list: List, ray_samples_list: List) -> Dict:
        """NerfactoModel._get_outputs, models/nerfacto.py:299-353."""
        field_outputs = field_.forward(ray_samples, compute_normals=self.config.predict_normals)
        weights = ray_samples.get_weights(field_outputs[FieldHeadNames.DENSITY])
        weights_list.append(weights)
        ray_samples_list.append(ray_samples)
        colour = field_outputs[FieldHeadNames.RGB]
        bg_mode, bg_const, bg_per_ray = renderer._bg_args(renderer.background_color, colour.shape[-1])
        if bg_per_ray is None and weights.dim() == 3:
            # the four renderer calls below (models/nerfacto.py:316-320) read the same weights: one launch
            r_, s_ = weights.shape[0], weights.shape[1]
            lay = getattr(ray_samples, "_layout", None)
            in_place = lay is not None and lay.ebins.shape == (r_, s_ + 1)
            rgb, accumulation, depth, exp_raw, minmax = ops.render(
                weights.reshape(r_, s_), colour, None if in_place else ray_samples.frustums.starts,
                None if in_place else ray_samples.frustums.ends, bg_mode=bg_mode, bg=bg_const,
                eval_mode=not self.training, want_depth=True, bins=lay.ebins if in_place else None)
            expected_depth = torch.clamp(exp_raw, minmax[0], minmax[1])  # batch-global clip, renderers.py:574
        else:
            rgb = renderer(rgb=colour, weights=weights)
            with torch.no_grad():
                depth = self.renderer_depth(weights=weights, ray_samples=ray_samples)
            expected_depth = self.renderer_expected_depth(weights=weights, ray_samples=ray_samples)
            accumulation = self.renderer_accumulation(weights=weights)
        outputs = {"rgb": rgb, "accumulation": accumulation, "depth": depth, "expected_depth": expected_depth,
                   "density": field_outputs[FieldHeadNames.DENSITY]}
        if self.training:
            outputs["weights_list"] = weights_list
            outputs["ray_samples_list"] = ray_samples_list
        with torch.no_grad():
            for i in range(self.config.num_proposal_iterations):
                outputs[f"prop_depth_{i}"] = self.renderer_depth(weights=weights_list[i],
                                                                 ray_samples=ray_samples_list[i])
        outputs["_field_rgb"] = field_outputs[FieldHeadNames.RGB]
        return outputs

    def _levels_fusable(self, ray_bundle: RayBundle, sampler: ProposalNetworkSampler, density_fns, renderer) -> bool:
        """The per-level launches cover the configuration thermal-nerfacto ships: piecewise initial sampler,
        PDFSampler without the original bins, proposal fields evaluated on the samplers' ray layout, a background
        that is the same for every ray, at most two proposal levels."""
        if not (self.fuse_levels and ray_bundle.origins.is_cuda and ray_bundle.origins.dim() == 2):
            return False
        if not (type(sampler.initial_sampler) is UniformLinDispPiecewiseSampler and type(sampler.pdf_sampler) is PDFSampler
                and not sampler.pdf_sampler.include_original and 1 <= sampler.num_proposal_network_iterations <= 2):
            return False
        if self.config.predict_normals or ray_bundle.nears is None or ray_bundle.fars is None:
            return False
        for fn in density_fns[:sampler.num_proposal_network_iterations]:
            owner = getattr(fn, "__self__", None)
            if not (isinstance(owner, HashMLPDensityField) and getattr(fn, "__name__", "") == "density_fn"):
                return False
        bg = renderer.background_color
        if isinstance(bg, Tensor) or (bg == "random" and self.training):
            return False
        return True

    def _branch_sample(self, ray_bundle: RayBundle, sampler: ProposalNetworkSampler, density_fns, renderer,
                       jitters: Optional[List[Tensor]]) -> "_BranchSamples":
        """Proposal phase of one branch: ProposalNetworkSampler.generate_ray_samples (ray_samplers.py:576-618).  With
        per-level launches (_levels_fusable) each level is its proposal field + ONE launch (weights, median depth,
        resampling: csrc/tn_level.cu); otherwise the sampler module runs component by component."""
        bs = _BranchSamples(ray_bundle, sampler)
        if not self._levels_fusable(ray_bundle, sampler, density_fns, renderer):
            bs.generic = sampler(ray_bundle, density_fns=density_fns, jitters=jitters)
            return bs
        n = sampler.num_proposal_network_iterations
        updated = sampler.will_update() if sampler._forced_updated is None else sampler._forced_updated
        dev = ray_bundle.origins.device
        num_rays = ray_bundle.origins.shape[0]
        nears, fars = ray_bundle.nears.reshape(-1), ray_bundle.fars.reshape(-1)
        if jitters is None:
            jitters = sampler.draw_jitters(num_rays, dev)
        jit = (lambda i: None) if jitters is None else (lambda i: jitters[i])
        bs.spacing = _PiecewiseSpacing(ray_bundle.nears, ray_bundle.fars)
        bs.counts = list(sampler.num_proposal_samples_per_ray[:n]) + [sampler.num_nerf_samples_per_ray]
        sbins, ebins = ops.piecewise_bins(nears, fars, bs.counts[0], jit(0))
        anneal = sampler._anneal_tensor(dev)
        for i in range(n):
            lay = RayLayout(ray_bundle.origins, ray_bundle.directions, ebins, sbins, nears, fars, chain=True)
            rs = samples_from_layout(ray_bundle, lay, bs.spacing)
            with torch.set_grad_enabled(updated and torch.is_grad_enabled()):
                sigma = density_fns[i].__self__.get_density(rs)[0].view(num_rays, bs.counts[i])
            # the proposal field handed the bundle on (pass-through outputs, see fields._chained_positions): the next
            # level / the main field read the chained tensors
            ray_bundle.origins, ray_bundle.directions = lay.origins, lay.directions
            w, med, sbins, ebins = fused_ops.level_resample(
                sigma.detach(), lay.ebins, lay.sbins, nears, fars, bs.counts[i + 1], jit(i + 1), anneal=anneal,
                histogram_padding=sampler.pdf_sampler.histogram_padding)
            bs.weights_list.append(w[..., None])
            bs.ray_samples_list.append(rs)
            bs.prop_sigma.append(sigma)
            bs.prop_depths.append(med)
        if sampler._forced_updated is None:
            sampler.mark_sampled(updated)
        bs.ebins, bs.sbins, bs.nears, bs.fars = ebins, sbins, nears, fars
        bs.grad_props = self.training and updated and torch.is_grad_enabled()
        return bs

    def _branch_render(self, bs: "_BranchSamples", field_, renderer) -> Tuple[Dict, RaySamples]:
        """Render phase of one branch: NerfactoModel._get_outputs (models/nerfacto.py:299-353) + the interlevel /
        distortion terms of get_loss_dict / get_metrics_dict (models/nerfacto.py:355-383).  With per-level launches:
        the field, then ONE launch for weights + all renderers + both losses (and one for the whole ray-level
        backward).  Returns (outputs of _get_outputs, final RaySamples)."""
        ray_bundle = bs.ray_bundle
        if bs.generic is not None:
            ray_samples, weights_list, ray_samples_list = bs.generic
            return self._get_outputs(ray_bundle, field_, renderer, ray_samples, weights_list, ray_samples_list), ray_samples
        n = len(bs.prop_sigma)
        num_rays = ray_bundle.origins.shape[0]
        lay = RayLayout(ray_bundle.origins, ray_bundle.directions, bs.ebins, bs.sbins, bs.nears, bs.fars, chain=True)
        ray_samples = samples_from_layout(ray_bundle, lay, bs.spacing)
        field_outputs = field_.forward(ray_samples, compute_normals=False)
        density, colour = field_outputs[FieldHeadNames.DENSITY], field_outputs[FieldHeadNames.RGB]
        bg_mode, bg_const, _ = renderer._bg_args(renderer.background_color, colour.shape[-1])
        rgb, accumulation, depth, expected_depth, _, weights, dist, inter = fused_ops.ray_heads(
            density.view(num_rays, bs.counts[n]), colour, bs.ebins, bs.sbins, bg_mode=bg_mode, bg=bg_const,
            eval_mode=not self.training, want_losses=self.training,
            prop_sigma=bs.prop_sigma if bs.grad_props else (),
            prop_ebins=[r_._layout.ebins for r_ in bs.ray_samples_list],
            prop_sbins=[r_._layout.sbins for r_ in bs.ray_samples_list],
            prop_weights=[w_[..., 0] for w_ in bs.weights_list],
            # training (a few thousand rays, launch-bound): the kernel's last CTA finishes extrema / loss means / clip.
            # Eval chunks (32768 rays = 8192 CTAs): one ticket per CTA and a one-CTA clip pass cost more (+40 us) than
            # the two small launches they replace, so the self-initialising path is kept there.
            scratch=fused_ops.heads_scratch(self._scratch, density.device, id(field_)) if self.training else None)
        weights_list = bs.weights_list + [weights[..., None]]
        ray_samples_list = bs.ray_samples_list + [ray_samples]
        outputs = {"rgb": rgb, "accumulation": accumulation, "depth": depth,
                   "expected_depth": expected_depth,  # clipped to the batch-global extrema (renderers.py:574) in-kernel
                   "density": density}
        if self.training:
            outputs["weights_list"] = weights_list
            outputs["ray_samples_list"] = ray_samples_list
            outputs["_distortion"], outputs["_interlevel"] = dist, inter
        for i in range(n):
            outputs[f"prop_depth_{i}"] = bs.prop_depths[i]
        outputs["_field_rgb"] = colour
        return outputs, ray_samples

    # The forward in two phases.  get_outputs runs them back to back; a data-parallel runner may put work between
    # them: nothing in the sample phase reads the main fields' parameters (engine.GraphedTrainStep, "pipeline").
    def sample_phase(self, ray_bundle: RayBundle, jitters: Optional[List[Tensor]] = None,
                     jitters_thermal: Optional[List[Tensor]] = None) -> "_ForwardState":
        """Pose corrections and proposal sampling of every branch (models/thermal_nerfacto.py:403-440): camera
        optimizers, proposal networks, resampling.  The RGB and the thermal branch are independent: with
        `branch_streams` the thermal one is issued on a second stream (see _render_phase)."""
        c = self.config
        # the reference deep-copies the bundle so that each path applies its own pose corrections (:407)
        ray_bundle_thermal = RayBundle(origins=ray_bundle.origins, directions=ray_bundle.directions,
                                       pixel_area=ray_bundle.pixel_area, camera_indices=ray_bundle.camera_indices,
                                       nears=ray_bundle.nears, fars=ray_bundle.fars, metadata=ray_bundle.metadata,
                                       times=ray_bundle.times)
        separate = c.density_mode == "separate"
        st = _ForwardState()
        st.streams = bool(separate and ray_bundle.origins.is_cuda and (
            (self.branch_streams and self.training) or (self.eval_branch_streams and not self.training)))
        renderer_rgb = self.renderer_rgbt if c.density_mode == "shared" else self.renderer_rgb
        dev = ray_bundle.origins.device
        if st.streams:
            # the jitter draws are made first, in the reference's order, so the random stream is unchanged
            num_rays = ray_bundle.origins.shape[0]
            if jitters is None:
                jitters = self.proposal_sampler.draw_jitters(num_rays, dev)
            if jitters_thermal is None:
                jitters_thermal = self.proposal_sampler_thermal.draw_jitters(num_rays, dev)
            main = torch.cuda.current_stream(dev)
            if self._side_stream is None or self._side_stream.device != dev:
                self._side_stream = torch.cuda.Stream(device=dev)
            side = self._side_stream
            side.wait_stream(main)
            with torch.cuda.stream(side):
                st.thermal = self._thermal_sample(ray_bundle_thermal, jitters_thermal)
        self.shared_camera_optimizer.apply_to_raybundle(ray_bundle)
        if self.training:
            self.camera_optimizer.apply_to_raybundle(ray_bundle)
        st.rgb = self._branch_sample(ray_bundle, self.proposal_sampler, self.density_fns, renderer_rgb, jitters)
        if separate and not st.streams:
            st.thermal = self._thermal_sample(ray_bundle_thermal, jitters_thermal)
        return st

    def _thermal_sample(self, ray_bundle_thermal: RayBundle, jitters_thermal) -> "_BranchSamples":
        """Pose corrections and proposal sampling of the thermal branch (:431-440)."""
        self.shared_camera_optimizer_thermal.apply_to_raybundle(ray_bundle_thermal)
        if self.training:
            self.camera_optimizer_thermal.apply_to_raybundle(ray_bundle_thermal)
        return self._branch_sample(ray_bundle_thermal, self.proposal_sampler_thermal, self.density_fns_thermal,
                                   self.renderer_thermal, jitters_thermal)

    def get_outputs(self, ray_bundle: RayBundle, jitters: Optional[List[Tensor]] = None,
                    jitters_thermal: Optional[List[Tensor]] = None) -> Dict:
        """ThermalNerfactoModel.get_outputs, models/thermal_nerfacto.py:403-489."""
        return self.render_phase(self.sample_phase(ray_bundle, jitters, jitters_thermal))

    def render_phase(self, st: "_ForwardState") -> Dict:
        """Main fields, renderers, cross-field terms and output assembly (models/thermal_nerfacto.py:441-489) for the
        samples of `sample_phase`.

        Training, density_mode="separate", CUDA: the RGB and the thermal branch are independent until the cross-field
        terms, which are independent of each other, and most of their kernels are latency- or issue-bound.  The
        thermal branch and one cross term are therefore issued on a second stream: in the captured train step they
        (and, through autograd's stream tracking, their backwards) become parallel arms of the graph.  Measured:
        2.93 -> 2.83 ms/step with the thermal branch on the second stream, -> 2.67 with the cross terms overlapped as
        well; a four-stream version (cross terms started right after the samplers) was slower (2.73)."""
        c = self.config
        separate = c.density_mode == "separate"
        want_cross = separate and (c.density_loss_mult > 0 or not self.training)
        renderer_rgb = self.renderer_rgbt if c.density_mode == "shared" else self.renderer_rgb
        thermal_outputs = ray_samples_thermal = cross = None
        if st.streams:
            dev = st.rgb.ray_bundle.origins.device
            main, side = torch.cuda.current_stream(dev), self._side_stream
            side.wait_stream(main)
            with torch.cuda.stream(side):
                thermal_outputs, ray_samples_thermal = self._branch_render(st.thermal, self.field_thermal,
                                                                           self.renderer_thermal)
            outputs, ray_samples = self._branch_render(st.rgb, self.field, renderer_rgb)
            main.wait_stream(side)
            if want_cross:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    d2t = self.field_thermal.get_density_only(ray_samples)
                d2 = self.field.get_density_only(ray_samples_thermal)
                main.wait_stream(side)
                cross = (d2, d2t)
            crossing = list(thermal_outputs.values()) + [ray_samples_thermal._layout.ebins,
                                                         ray_samples_thermal._layout.sbins]
            crossing += [cross[1]] if cross is not None else []
            for v in crossing:  # made on the side stream, consumed (and eventually freed) on the main one
                if torch.is_tensor(v):
                    v.record_stream(main)
        else:
            outputs, ray_samples = self._branch_render(st.rgb, self.field, renderer_rgb)
            if separate:
                thermal_outputs, ray_samples_thermal = self._branch_render(st.thermal, self.field_thermal,
                                                                           self.renderer_thermal)
                if want_cross:
                    # cross-field densities for the density regulariser (:447-458).  The reference runs the full
                    # field forward here and throws the colour away; only the density is evaluated.
                    cross = (self.field.get_density_only(ray_samples_thermal),
                             self.field_thermal.get_density_only(ray_samples))
        field_rgb = outputs.pop("_field_rgb")
        for rs_ in (ray_samples, ray_samples_thermal, *st.rgb.ray_samples_list,
                    *(st.thermal.ray_samples_list if st.thermal is not None else ())):
            lay_ = getattr(rs_, "_layout", None)
            if lay_ is not None:  # this forward's consumers are all in place: later, independent evaluations of the
                lay_.chain = False  # returned samples must not hook into its gradient chain (rays.RayLayout.chain)

        if c.density_mode == "shared":
            rgbt = outputs["rgb"]
            outputs["rgbt"] = rgbt
            outputs["rgb"] = rgbt[..., :3]
            outputs["rgb_thermal"] = rgbt[..., 3:]
        elif separate:
            field_rgb_thermal = thermal_outputs.pop("_field_rgb")
            for k, v in thermal_outputs.items():
                outputs[f"{k}_thermal"] = v
            if cross is not None:
                outputs["density2"], outputs["density2_thermal"] = cross

            if not self.training:
                # "removal" renders (:460-487).  The reference recomputes both field forwards on the very same
                # samples (identical values in eval); the colours from above are reused instead.
                thr = c.removal_min_density_diff
                mask_rgb = (outputs["density"] / outputs["density"]
                            - outputs["density2_thermal"] / outputs["density"]).abs() < thr
                w_rm = ray_samples.get_weights(outputs["density"] * mask_rgb)
                outputs["removal"] = self.renderer_rgb(rgb=field_rgb, weights=w_rm)
                mask_th = (outputs["density_thermal"] / outputs["density_thermal"]
                           - outputs["density2"] / outputs["density_thermal"]).abs() < thr
                # reference quirk kept: thermal densities composited with the RGB samples' deltas (:485)
                w_rm_th = ray_samples.get_weights(outputs["density_thermal"] * mask_th)
                outputs["removal_thermal"] = self.renderer_thermal(rgb=field_rgb_thermal, weights=w_rm_th)
        return outputs

    @torch.no_grad()
    def get_outputs_for_camera_ray_bundle(self, camera_ray_bundle: RayBundle) -> Dict[str, Tensor]:
        """models/base_model.py:177-206: chunked full-frame render (chunk size kept, the expected-depth clip is
        chunk-global in the reference)."""
        input_device = camera_ray_bundle.directions.device
        chunk = self.config.eval_num_rays_per_chunk
        image_height, image_width = camera_ray_bundle.origins.shape[:2]
        num_rays = len(camera_ray_bundle)
        flat = camera_ray_bundle.flatten()
        outputs_lists = defaultdict(list)
        for i in range(0, num_rays, chunk):
            ray_bundle = flat[i:i + chunk].to(self.device)
            outputs = self.forward(ray_bundle=ray_bundle)
            for name, out in outputs.items():
                if torch.is_tensor(out):
                    outputs_lists[name].append(out.to(input_device))
        return {name: torch.cat(lst).view(image_height, image_width, -1) for name, lst in outputs_lists.items()}

    # -------------------------------------------------------------------------------------- losses
    def get_metrics_dict(self, outputs, batch) -> Dict:
        """models/thermal_nerfacto.py:253-282 without the PSNR entries (torchmetrics is the reference's
        reporting code, not part of the loss)."""
        metrics_dict = {}
        if self.training:
            total = None  # the reference starts from 0 and adds (:255-262): 0 + a == a, one launch less
            for s in self.output_suffixes:
                fused = outputs.get(f"_distortion{s}")  # made by the final level's launch (_branch_fused)
                term = fused if fused is not None else distortion_loss(
                    outputs[f"weights_list{s}"], outputs[f"ray_samples_list{s}"])
                total = term if total is None else total + term
            metrics_dict["distortion"] = total
        self.camera_optimizer.get_metrics_dict(metrics_dict)
        self.shared_camera_optimizer.get_metrics_dict(metrics_dict)
        if self.config.density_mode == "separate":
            self.camera_optimizer_thermal.get_metrics_dict(metrics_dict)
            self.shared_camera_optimizer_thermal.get_metrics_dict(metrics_dict)
        return metrics_dict

    def get_loss_dict(self, outputs, batch, metrics_dict=None) -> Dict[str, Tensor]:
        """models/thermal_nerfacto.py:284-388.  Every entry is (config weight) x (term); the terms are collected
        first and weighted + totalled by one launch (`fused_ops.loss_sum`) -- the returned dict carries the sum
        the trainer would form (engine/trainer.py:479) as `.total`."""
        c = self.config
        entries: List[Tuple[str, Tensor, float]] = []  # (key, un-weighted term, weight); keys may repeat
        image = batch["image"].to(self.device)
        is_thermal = batch["is_thermal"].to(self.device)
        fused_pixels = (self.fuse_losses and c.background_color != "random" and image.shape[-1] == 3
                        and image.shape[0] % 4 == 0)
        if fused_pixels:
            # rgb / thermal MSE, tv_pixel and cross_channel terms (:315-354) in one launch each way
            thermal = outputs["rgb_thermal"] if c.density_mode != "rgb_only" else None
            pl = fused_ops.pixel_losses(outputs["rgb"], thermal, image, is_thermal)
            entries.append(("rgb_loss", pl[0], 1.0))
            if c.density_mode != "rgb_only":
                entries.append(("thermal_loss", pl[1], c.thermal_loss_mult))
        else:
            if c.density_mode != "rgb_only":
                pred = torch.cat((outputs["rgb"], outputs["rgb_thermal"]), dim=1)
            else:
                pred = torch.cat((outputs["rgb"], torch.zeros(outputs["rgb"].shape[0], 1, device=self.device)), dim=1)
            pred_rgb, gt_rgb = self.renderer_rgbt.blend_background_for_loss_computation(
                pred_image=pred, pred_accumulation=outputs["accumulation"], gt_image=image, is_thermal=is_thermal)
            is_rgb = (1 - is_thermal)[:, None]
            entries.append(("rgb_loss", self.rgb_loss(gt_rgb[..., :3] * is_rgb, pred_rgb[..., :3] * is_rgb), 1.0))
            if c.density_mode != "rgb_only":
                th = is_thermal[:, None]
                entries.append(("thermal_loss", self.rgb_loss(gt_rgb[..., 3:] * th, pred_rgb[..., 3:] * th),
                                c.thermal_loss_mult))
        if c.density_mode == "separate" and c.density_loss_mult > 0:
            m, r = c.density_loss_mult, c.rgb_density_loss_mult
            d, d2, dt, d2t = (outputs["density"], outputs["density2"], outputs["density_thermal"],
                              outputs["density2_thermal"])
            if self.fuse_losses:  # the four L1 terms and their stop-gradient pattern (:328-344) in one launch
                entries.append(("density_loss", fused_ops.density_l1(d, d2, dt, d2t, m, r, scratch=self._scratch), 1.0))
            elif r == 1:
                entries.append(("density_loss", self.density_loss(d2, dt) + self.density_loss(d, d2t), m))
            else:  # asymmetric stop-gradient pattern, :336-344
                entries.append(("density_loss", self.density_loss(d2.detach(), dt) + self.density_loss(d.detach(), d2t)
                                + r * self.density_loss(d2, dt.detach()) + r * self.density_loss(d, d2t.detach()), m))
        if c.density_mode != "rgb_only" and c.tv_pixel_loss_mult > 0:
            entries.append(("tv_pixel_loss", pl[2] if fused_pixels else tv_pixel_loss(pred_rgb[..., 3:], is_thermal),
                            c.tv_pixel_loss_mult))
        if c.density_mode != "rgb_only" and c.cross_channel_loss_mult > 0:
            entries.append(("cross_channel_loss", pl[3] if fused_pixels
                            else cross_channel_loss(pred_rgb[..., 3:], gt_rgb[..., :3], is_thermal),
                            c.cross_channel_loss_mult))
        if self.training:
            assert metrics_dict is not None and "distortion" in metrics_dict
            for s in self.output_suffixes:
                fused = outputs.get(f"_interlevel{s}")  # made by the final level's launch (_branch_fused)
                entries.append(("interlevel_loss", fused if fused is not None else interlevel_loss(
                    outputs[f"weights_list{s}"], outputs[f"ray_samples_list{s}"]), c.interlevel_loss_mult))
            for s in self.output_suffixes:
                # reference quirk kept: the SUMMED distortion metric is added once per suffix (:368)
                entries.append(("distortion_loss", metrics_dict["distortion"], c.distortion_loss_mult))
        cams = [self.camera_optimizer] if self.training else []
        if self.training and c.density_mode == "separate":
            cams.append(self.camera_optimizer_thermal)
        cams.append(self.shared_camera_optimizer)
        if c.density_mode == "separate":
            cams.append(self.shared_camera_optimizer_thermal)
        for cam in cams:
            tmp: Dict[str, Tensor] = {}
            cam.get_loss_dict(tmp)
            entries.extend((k, v, 1.0) for k, v in tmp.items())
        return _assemble_losses(entries, self.fuse_losses)

    def get_train_loss_dict(self, ray_bundle: RayBundle, batch: Dict[str, Tensor], **fw):
        """VanillaPipeline.get_train_loss_dict body, pipelines/base_pipeline.py:291-304."""
        outputs = self(ray_bundle, **fw)
        metrics = self.get_metrics_dict(outputs, batch)
        return outputs, self.get_loss_dict(outputs, batch, metrics), metrics
