"""Oracle: MLPs, activations, contraction, SH basis and the nerfacto / proposal fields.

TEST INFRASTRUCTURE ONLY.  Weights are addressed by the reference's own state_dict keys so a
reference checkpoint can be fed in unchanged (SURVEY.md 8b "state_dict contract").
"""
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .hashgrid import hash_encode, hash_scalings


class _TruncExpFn(torch.autograd.Function):
    """field_components/activations.py:28-42: exp forward, backward uses exp(clamp(x,-15,15))."""

    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _TruncExpFn.apply


def mlp_forward(x: torch.Tensor, weights: List[torch.Tensor], biases: List[torch.Tensor],
                out_activation: Optional[str] = None) -> torch.Tensor:
    """field_components/mlp.py:159-178 with activation=ReLU, no skip connections."""
    n = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        x = F.linear(x, w, b)
        if i < n - 1:
            x = torch.relu(x)
    if out_activation == "sigmoid":
        x = torch.sigmoid(x)
    elif out_activation is not None:
        raise ValueError(out_activation)
    return x


def scene_contraction_linf(x: torch.Tensor) -> torch.Tensor:
    """field_components/spatial_distortions.py:66-69 with order=inf."""
    mag = torch.linalg.norm(x, ord=float("inf"), dim=-1)[..., None]
    return torch.where(mag < 1, x, (2 - (1 / mag)) * (x / mag))


def normalize_positions(positions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fields/nerfacto_field.py:207-215 (contraction branch): returns (x in [0,1] zeroed outside, selector)."""
    p = scene_contraction_linf(positions)
    p = (p + 2.0) / 4.0
    selector = ((p > 0.0) & (p < 1.0)).all(dim=-1)
    p = p * selector[..., None]
    return p, selector


def sh4_basis(d: torch.Tensor) -> torch.Tensor:
    """utils/math.py:29-95 with levels=4 (16 components).  The field evaluates it on (dir+1)/2
    without mapping back to [-1,1] and under no_grad (encodings.py:792-795)."""
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    xx, yy, zz = x**2, y**2, z**2
    c = torch.zeros((*d.shape[:-1], 16), device=d.device)
    c[..., 0] = 0.28209479177387814
    c[..., 1] = 0.4886025119029199 * y
    c[..., 2] = 0.4886025119029199 * z
    c[..., 3] = 0.4886025119029199 * x
    c[..., 4] = 1.0925484305920792 * x * y
    c[..., 5] = 1.0925484305920792 * y * z
    c[..., 6] = 0.9461746957575601 * zz - 0.31539156525251999
    c[..., 7] = 1.0925484305920792 * x * z
    c[..., 8] = 0.5462742152960396 * (xx - yy)
    c[..., 9] = 0.5900435899266435 * y * (3 * xx - yy)
    c[..., 10] = 2.890611442640554 * x * y * z
    c[..., 11] = 0.4570457994644658 * y * (5 * zz - 1)
    c[..., 12] = 0.3731763325901154 * z * (5 * zz - 3)
    c[..., 13] = 0.4570457994644658 * x * (5 * zz - 1)
    c[..., 14] = 1.445305721320277 * z * (xx - yy)
    c[..., 15] = 0.5900435899266435 * x * (xx - 3 * yy)
    return c


def _linear_stack(sd: Dict[str, torch.Tensor], prefix: str):
    ws, bs = [], []
    i = 0
    while f"{prefix}.layers.{i}.weight" in sd:
        ws.append(sd[f"{prefix}.layers.{i}.weight"])
        bs.append(sd[f"{prefix}.layers.{i}.bias"])
        i += 1
    return ws, bs


def density_field(sd: Dict[str, torch.Tensor], prefix: str, positions: torch.Tensor, *, num_levels: int,
                  base_res: int, max_res: int, log2_hashmap_size: int, average_init_density: float = 1.0):
    """NerfactoField.get_density, fields/nerfacto_field.py:205-229.

    positions [*bs,3] world space -> (density [*bs,1], geo features [*bs,15]).
    """
    x, selector = normalize_positions(positions)
    scal = hash_scalings(num_levels, base_res, max_res).to(x.device)
    enc = hash_encode(x.view(-1, 3), sd[f"{prefix}.mlp_base.model.0.hash_table"], scal, log2_hashmap_size)
    ws, bs = _linear_stack(sd, f"{prefix}.mlp_base.model.1")
    h = mlp_forward(enc, ws, bs).view(*positions.shape[:-1], -1)
    d_raw, geo = torch.split(h, [1, h.shape[-1] - 1], dim=-1)
    density = average_init_density * trunc_exp(d_raw.to(x))
    density = density * selector[..., None]
    return density, geo


def proposal_density(sd: Dict[str, torch.Tensor], prefix: str, positions: torch.Tensor, *, num_levels: int,
                     base_res: int, max_res: int, log2_hashmap_size: int, average_init_density: float = 1.0):
    """HashMLPDensityField.get_density via Field.density_fn, fields/density_fields.py:95-118,
    fields/base_field.py:48-69."""
    x, selector = normalize_positions(positions)
    scal = hash_scalings(num_levels, base_res, max_res).to(x.device)
    enc = hash_encode(x.view(-1, 3), sd[f"{prefix}.mlp_base.0.hash_table"], scal, log2_hashmap_size)
    ws, bs = _linear_stack(sd, f"{prefix}.mlp_base.1")
    raw = mlp_forward(enc, ws, bs).view(*positions.shape[:-1], -1).to(x)
    density = average_init_density * trunc_exp(raw)
    return density * selector[..., None]


def colour_head(sd: Dict[str, torch.Tensor], prefix: str, directions: torch.Tensor, geo: torch.Tensor,
                camera_indices: torch.Tensor, *, training: bool, use_average_appearance_embedding: bool = True):
    """NerfactoField.get_outputs, fields/nerfacto_field.py:272-348 (ThermalNerfactoField only changes
    the head's out_dim, fields/thermal_nerfacto_field.py:91-99).

    directions [R,S,3], geo [R,S,15], camera_indices [R,S,1] -> colour [R,S,C] after sigmoid.
    """
    with torch.no_grad():
        d = sh4_basis(((directions + 1.0) / 2.0).view(-1, 3))
    emb_w = sd[f"{prefix}.embedding_appearance.embedding.weight"]
    if training:
        emb = emb_w[camera_indices.squeeze()]
    elif use_average_appearance_embedding:
        emb = torch.ones((*directions.shape[:-1], emb_w.shape[1]), device=emb_w.device) * emb_w.mean(0)
    else:
        emb = torch.zeros((*directions.shape[:-1], emb_w.shape[1]), device=emb_w.device)
    h = torch.cat([d, geo.reshape(-1, geo.shape[-1]), emb.reshape(-1, emb_w.shape[1])], dim=-1)
    ws, bs = _linear_stack(sd, f"{prefix}.mlp_head")
    out = mlp_forward(h, ws, bs, out_activation="sigmoid")
    return out.view(*directions.shape[:-1], -1)
