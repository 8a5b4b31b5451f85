"""Fields: density + outputs, with the reference's class names, signatures and state_dict keys.

Mirrors `nerfstudio/fields/{base_field,nerfacto_field,thermal_nerfacto_field,density_fields}.py`.
"""
from enum import Enum
from typing import Dict, Literal, Optional, Tuple

import torch
from torch import Tensor, nn

from . import fused_ops, ops
from .field_components import (
    MLP,
    Embedding,
    HashEncoding,
    MLPWithHashEncoding,
    SceneContraction,
    SHEncoding,
    SpatialDistortion,
    trunc_exp,
)
from .rays import Frustums, RaySamples


class FieldHeadNames(Enum):
    """field_components/field_heads.py (names the path uses)."""

    RGB = "rgb"
    DENSITY = "density"
    NORMALS = "normals"
    PRED_NORMALS = "pred_normals"


def get_normalized_directions(directions: Tensor) -> Tensor:
    """fields/base_field.py:136-142."""
    return (directions + 1.0) / 2.0


def _chained_positions(lay) -> Tuple[Tensor, Tensor]:
    """ops.sample_positions for a RayLayout.  The layout's origins/directions are replaced by the call's pass-through
    outputs, so the bundle's consumers form a chain (proposal levels -> main field -> cross-field term) along which
    ONE gradient buffer is handed back and added onto in place (ops._SamplePositionsFn)."""
    if not _chainable(lay):  # outputs made under no_grad carry no graph: the layout must keep the originals
        return ops.sample_positions(lay.origins, lay.directions, lay.ebins)
    x, selector, lay.origins, lay.directions = ops.sample_positions(lay.origins, lay.directions, lay.ebins, chain=True)
    return x, selector


def _chainable(lay) -> bool:
    return lay.chain and torch.is_grad_enabled() and (lay.origins.requires_grad or lay.directions.requires_grad)


def _is_linf_contraction(sd: Optional[SpatialDistortion]) -> bool:
    return isinstance(sd, SceneContraction) and sd.order == float("inf")


class Field(nn.Module):
    """fields/base_field.py:40-133."""

    def __init__(self) -> None:
        super().__init__()
        self._sample_locations = None
        self.grads_ready_callback = None  # see NerfactoField.forward
        self._scratch = fused_ops.LaunchScratch()  # persistent, self-cleaning launch buffers (fused_ops.LaunchScratch)
        self._density_before_activation = None

    def _grid_coordinates(self, ray_samples: RaySamples) -> Tuple[Tensor, Tensor]:
        """(x [N,3] in [0,1] zeroed outside the box, selector [N]) for the samples; the steps of
        fields/nerfacto_field.py:207-215 == fields/density_fields.py:96-103."""
        lay = getattr(ray_samples, "_layout", None)  # None for reference-built RaySamples
        if _is_linf_contraction(self.spatial_distortion):
            if lay is not None:
                return _chained_positions(lay)
            return ops.contract_points(ray_samples.frustums.get_positions().reshape(-1, 3))
        positions = ray_samples.frustums.get_positions()
        if self.spatial_distortion is not None:
            positions = (self.spatial_distortion(positions) + 2.0) / 4.0
        else:  # SceneBox.get_normalized_positions, data/scene_box.py
            positions = (positions - self.aabb[0]) / (self.aabb[1] - self.aabb[0])
        selector = ((positions > 0.0) & (positions < 1.0)).all(dim=-1)
        positions = positions * selector[..., None]
        return positions.reshape(-1, 3), selector.reshape(-1).to(positions.dtype)

    def density_fn(self, positions: Tensor, times: Optional[Tensor] = None) -> Tensor:
        """fields/base_field.py:48-69."""
        del times
        ray_samples = RaySamples(
            frustums=Frustums(origins=positions, directions=torch.ones_like(positions),
                              starts=torch.zeros_like(positions[..., :1]), ends=torch.zeros_like(positions[..., :1]),
                              pixel_area=torch.ones_like(positions[..., :1])))
        density, _ = self.get_density(ray_samples)
        return density

    def get_density(self, ray_samples: RaySamples):
        raise NotImplementedError

    def get_outputs(self, ray_samples: RaySamples, density_embedding: Optional[Tensor] = None):
        raise NotImplementedError

    def forward(self, ray_samples: RaySamples, compute_normals: bool = False) -> Dict[FieldHeadNames, Tensor]:
        """fields/base_field.py:114-133."""
        if compute_normals:
            raise NotImplementedError("analytic normals (predict_normals) are outside the thermal-nerfacto hot path")
        density, density_embedding = self.get_density(ray_samples)
        field_outputs = self.get_outputs(ray_samples, density_embedding=density_embedding)
        field_outputs[FieldHeadNames.DENSITY] = density
        return field_outputs


class NerfactoField(Field):
    """fields/nerfacto_field.py:47-348 (the north star's `TCNNNerfactoField` is this class's legacy name).

    Unsupported reference options raise at construction: transient embedding, semantics, predicted normals.
    """

    aabb: Tensor

    def __init__(self, aabb: Tensor, num_images: int, num_layers: int = 2, hidden_dim: int = 64, geo_feat_dim: int = 15,
                 num_levels: int = 16, base_res: int = 16, max_res: int = 2048, log2_hashmap_size: int = 19,
                 num_layers_color: int = 3, num_layers_transient: int = 2, features_per_level: int = 2,
                 hidden_dim_color: int = 64, hidden_dim_transient: int = 64, appearance_embedding_dim: int = 32,
                 transient_embedding_dim: int = 16, use_transient_embedding: bool = False, use_semantics: bool = False,
                 num_semantic_classes: int = 100, pass_semantic_gradients: bool = False, use_pred_normals: bool = False,
                 use_average_appearance_embedding: bool = False,
                 spatial_distortion: Optional[SpatialDistortion] = None, average_init_density: float = 1.0,
                 implementation: Literal["tcnn", "torch", "b200"] = "b200", num_channels: int = 3) -> None:
        super().__init__()
        if use_transient_embedding or use_semantics or use_pred_normals:
            raise NotImplementedError("transient/semantic/pred-normal heads are not part of thermal-nerfacto")
        self.register_buffer("aabb", aabb)
        self.geo_feat_dim = geo_feat_dim
        self.register_buffer("max_res", torch.tensor(max_res))
        self.register_buffer("num_levels", torch.tensor(num_levels))
        self.register_buffer("log2_hashmap_size", torch.tensor(log2_hashmap_size))
        self.spatial_distortion = spatial_distortion
        self.num_images = num_images
        self.appearance_embedding_dim = appearance_embedding_dim
        self.embedding_appearance = Embedding(num_images, appearance_embedding_dim) if appearance_embedding_dim > 0 \
            else None
        self.use_average_appearance_embedding = use_average_appearance_embedding
        self.use_transient_embedding = self.use_semantics = self.use_pred_normals = False
        self.base_res = base_res
        self.average_init_density = average_init_density
        self.step = 0
        self.direction_encoding = SHEncoding(levels=4, implementation=implementation)
        self.mlp_base = MLPWithHashEncoding(
            num_levels=num_levels, min_res=base_res, max_res=max_res, log2_hashmap_size=log2_hashmap_size,
            features_per_level=features_per_level, num_layers=num_layers, layer_width=hidden_dim,
            out_dim=1 + geo_feat_dim, activation=nn.ReLU(), out_activation=None, implementation=implementation)
        self.mlp_head = MLP(
            in_dim=self.direction_encoding.get_out_dim() + geo_feat_dim + appearance_embedding_dim,
            num_layers=num_layers_color, layer_width=hidden_dim_color, out_dim=num_channels, activation=nn.ReLU(),
            out_activation=nn.Sigmoid(), implementation=implementation)

    def get_density(self, ray_samples: RaySamples) -> Tuple[Tensor, Tensor]:
        """fields/nerfacto_field.py:205-229."""
        x, selector = self._grid_coordinates(ray_samples)
        shape = ray_samples.frustums.shape
        self._sample_locations = x.view(*shape, 3)
        h = self.mlp_base(self._sample_locations)  # [R,S,3] tells the encode kernels the ray structure
        h_flat = h.view(-1, h.shape[-1])
        density_before_activation, base_mlp_out = torch.split(h, [1, self.geo_feat_dim], dim=-1)
        self._density_before_activation = density_before_activation
        # average_init_density * trunc_exp(h0) * selector in one kernel (:227-228)
        density = fused_ops.density_act(h_flat, selector, self.average_init_density).view(*shape, 1)
        return density, base_mlp_out

    def get_density_only(self, ray_samples: RaySamples) -> Tensor:
        """get_density(...)[0] without the geometry features: only row 0 of the density MLP's last layer is evaluated
        and `average_init_density * trunc_exp(.) * selector` (:227-228) is that kernel's epilogue.  Used for the
        cross-field density terms (models/thermal_nerfacto.py:447-458), which discard everything else."""
        mlp = self.mlp_base.model[1]
        if mlp._out_act != ops.ACT_NONE or len(mlp.layers) < 2:
            return self.get_density(ray_samples)[0]
        x, selector = self._grid_coordinates(ray_samples)
        feats = self.mlp_base.model[0](x.view(*ray_samples.frustums.shape, 3)).view(x.shape[0], -1)
        ws = [l.weight for l in mlp.layers[:-1]] + [mlp.layers[-1].weight[:1]]
        bs = [l.bias for l in mlp.layers[:-1]] + [mlp.layers[-1].bias[:1]]
        sinks = None
        if mlp.grad_sinks is not None:
            lw, lb = mlp.grad_sinks[-1]
            sinks = [*mlp.grad_sinks[:-1], (lw[:1], lb[:1])]
        density = ops.mlp(feats, ws, bs, ops.ACT_TRUNC_EXP, sinks=sinks, row_mul=selector,
                          out_scale=self.average_init_density)
        return density.view(*ray_samples.frustums.shape, 1)

    def forward(self, ray_samples: RaySamples, compute_normals: bool = False) -> Dict[FieldHeadNames, Tensor]:
        """Field.forward (fields/base_field.py:114-133).  For samples that carry a per-ray layout the density
        activation, the geo/SH/appearance concatenation (:335-344) and its backward run as one kernel each
        (`fused_ops.field_split`); the values are those of get_density + get_outputs."""
        lay = getattr(ray_samples, "_layout", None)  # None for reference-built RaySamples
        if compute_normals or lay is None or not _is_linf_contraction(self.spatial_distortion) \
                or ray_samples.camera_indices is None:
            return super().forward(ray_samples, compute_normals=compute_normals)
        rays, samples = lay.num_rays, lay.num_samples
        x, selector = _chained_positions(lay)
        if self.grads_ready_callback is not None and torch.is_grad_enabled() and x.requires_grad:
            # dL/dx of the main evaluation is produced by the encode backward, the LAST kernel that touches this
            # field's parameters in a step (heads and density MLP run before it, the cross-field density term, created
            # later in the forward, has run earlier): data-parallel runners start this field's all-reduce from here
            cb = self.grads_ready_callback
            x.register_hook(lambda _g: (cb(), None)[1])
        self._sample_locations = x.view(rays, samples, 3)
        h = self.mlp_base(self._sample_locations).view(rays * samples, -1)
        self._density_before_activation = h.view(rays, samples, -1)[..., :1]
        head = self.mlp_head
        emb = self.embedding_appearance
        if emb is not None and self.training and self.direction_encoding.levels == 4:
            # SH basis + the rays' embedding rows in one launch (:284-290: the samples of a ray share the camera); the
            # embedding's gradient is formed by the head's backward (fused_ops._FieldHeadFn)
            cam = ray_samples.camera_indices[:, 0, 0]
            sh, emb_ray = fused_ops.ray_features(lay.directions, emb.embedding.weight, cam)
            if fused_ops.field_head_supported(h.shape[-1], self.geo_feat_dim, emb_ray.shape[-1], samples, head):
                density, rgb = fused_ops.field_head(
                    h, selector, sh, emb_ray, rays, samples, self.geo_feat_dim, self.average_init_density,
                    [l.weight for l in head.layers], [l.bias for l in head.layers], head._out_act,
                    sinks=head.grad_sinks, emb_weight=emb.embedding.weight, cam_idx=cam, emb_sink=emb.grad_sink,
                    scratch=self._scratch)
                return {FieldHeadNames.RGB: rgb.view(rays, samples, -1),
                        FieldHeadNames.DENSITY: density.view(rays, samples, 1)}
            emb_ray = fused_ops.embed_rows(emb.embedding.weight, cam)  # differentiable lookup for the generic head
        else:
            sh = self.direction_encoding(get_normalized_directions(lay.directions))
            emb_ray = None
            if emb is not None:
                if self.training:
                    emb_ray = fused_ops.embed_rows(emb.embedding.weight, ray_samples.camera_indices[:, 0, 0])
                elif self.use_average_appearance_embedding:
                    emb_ray = emb.mean(dim=0)[None, :].expand(rays, -1)
                else:
                    emb_ray = torch.zeros((rays, self.appearance_embedding_dim), device=x.device)
        if emb_ray is not None and fused_ops.field_head_supported(h.shape[-1], self.geo_feat_dim, emb_ray.shape[-1],
                                                                  samples, head):
            # split + concatenation + colour head as one autograd node (one backward kernel)
            density, rgb = fused_ops.field_head(h, selector, sh, emb_ray, rays, samples, self.geo_feat_dim,
                                                self.average_init_density, [l.weight for l in head.layers],
                                                [l.bias for l in head.layers], head._out_act, sinks=head.grad_sinks)
            rgb = rgb.view(rays, samples, -1)
        else:
            density, head_in = fused_ops.field_split(h, selector, sh, emb_ray, rays, samples, self.geo_feat_dim,
                                                     self.average_init_density)
            rgb = head(head_in).view(rays, samples, -1)
        return {FieldHeadNames.RGB: rgb, FieldHeadNames.DENSITY: density.view(rays, samples, 1)}

    def get_outputs(self, ray_samples: RaySamples, density_embedding: Optional[Tensor] = None):
        """fields/nerfacto_field.py:272-348."""
        assert density_embedding is not None
        if ray_samples.camera_indices is None:
            raise AttributeError("Camera indices are not provided.")
        shape = ray_samples.frustums.directions.shape[:-1]
        lay = getattr(ray_samples, "_layout", None)  # None for reference-built RaySamples
        if lay is not None:  # SH once per ray, broadcast over the samples
            d = self.direction_encoding(get_normalized_directions(lay.directions))
            d = d[:, None, :].expand(*shape, d.shape[-1])
        else:
            d = self.direction_encoding(get_normalized_directions(ray_samples.frustums.directions).reshape(-1, 3))
            d = d.view(*shape, -1)
        parts = [d.reshape(-1, d.shape[-1]), density_embedding.reshape(-1, self.geo_feat_dim)]
        if self.embedding_appearance is not None:
            if self.training:
                emb = self.embedding_appearance(ray_samples.camera_indices.squeeze(-1))
            elif self.use_average_appearance_embedding:
                emb = torch.ones((*shape, self.appearance_embedding_dim), device=d.device) \
                    * self.embedding_appearance.mean(dim=0)
            else:
                emb = torch.zeros((*shape, self.appearance_embedding_dim), device=d.device)
            parts.append(emb.reshape(-1, self.appearance_embedding_dim))
        h = torch.cat(parts, dim=-1)
        rgb = self.mlp_head(h).view(*shape, -1)
        return {FieldHeadNames.RGB: rgb}


class ThermalNerfactoField(NerfactoField):
    """fields/thermal_nerfacto_field.py:10-99: NerfactoField whose colour head has `num_channels` outputs
    (3 RGB, 4 RGBT shared, 1 thermal)."""

    def __init__(self, aabb: Tensor, num_images: int, num_channels: int = 4, **kwargs) -> None:
        super().__init__(aabb, num_images, num_channels=num_channels, **kwargs)


class HashMLPDensityField(Field):
    """fields/density_fields.py:34-121.  `encoding.hash_table` and `mlp_base.0.hash_table` are the same
    Parameter under two state_dict keys, as in the reference."""

    aabb: Tensor

    def __init__(self, aabb: Tensor, num_layers: int = 2, hidden_dim: int = 64,
                 spatial_distortion: Optional[SpatialDistortion] = None, use_linear: bool = False, num_levels: int = 8,
                 max_res: int = 1024, base_res: int = 16, log2_hashmap_size: int = 18, features_per_level: int = 2,
                 average_init_density: float = 1.0, implementation: Literal["tcnn", "torch", "b200"] = "b200") -> None:
        super().__init__()
        self.register_buffer("aabb", aabb)
        self.spatial_distortion = spatial_distortion
        self.use_linear = use_linear
        self.fuse = True  # use the single-kernel path for ray samples when the shapes allow (see get_density)
        self.average_init_density = average_init_density
        self.register_buffer("max_res", torch.tensor(max_res))
        self.register_buffer("num_levels", torch.tensor(num_levels))
        self.register_buffer("log2_hashmap_size", torch.tensor(log2_hashmap_size))
        self.encoding = HashEncoding(num_levels=num_levels, min_res=base_res, max_res=max_res,
                                     log2_hashmap_size=log2_hashmap_size, features_per_level=features_per_level,
                                     implementation=implementation)
        if not self.use_linear:
            network = MLP(in_dim=self.encoding.get_out_dim(), num_layers=num_layers, layer_width=hidden_dim, out_dim=1,
                          activation=nn.ReLU(), out_activation=None, implementation=implementation)
            self.mlp_base = torch.nn.Sequential(self.encoding, network)
        else:
            self.linear = torch.nn.Linear(self.encoding.get_out_dim(), 1)

    def get_density(self, ray_samples: RaySamples) -> Tuple[Tensor, None]:
        """fields/density_fields.py:95-118."""
        lay = getattr(ray_samples, "_layout", None)  # None for reference-built RaySamples
        enc = self.encoding
        if (self.fuse and lay is not None and _is_linf_contraction(self.spatial_distortion) and not self.use_linear
                and enc.features_per_level == 2 and enc.num_levels <= 8 and not enc.use_half_table
                and len(self.mlp_base[1].layers) == 2 and self.mlp_base[1].layer_width == 16):
            # whole field in one kernel: positions -> contraction -> hash grid -> 16-wide MLP -> trunc_exp * selector
            l0, l1 = self.mlp_base[1].layers
            args = (lay.origins, lay.directions, lay.ebins, enc.hash_table, l0.weight, l0.bias, l1.weight, l1.bias,
                    enc.spec, self.average_init_density, enc.grad_sink, self.mlp_base[1].grad_sinks)
            if _chainable(lay):
                density, lay.origins, lay.directions = fused_ops.prop_density(*args, chain=True)
            else:
                density = fused_ops.prop_density(*args)
            return density.view(lay.num_rays, lay.num_samples, 1), None
        x, selector = self._grid_coordinates(ray_samples)
        shape = ray_samples.frustums.shape
        xs = x.view(*shape, 3)
        if not self.use_linear:
            raw = self.mlp_base(xs).view(x.shape[0], -1)
        else:
            raw = self.linear(self.encoding(xs)).view(x.shape[0], -1)
        # average_init_density * trunc_exp(raw) * selector in one kernel (:116-117)
        density = fused_ops.density_act(raw, selector, self.average_init_density).view(*shape, 1)
        return density, None

    def get_outputs(self, ray_samples: RaySamples, density_embedding: Optional[Tensor] = None) -> dict:
        return {}
