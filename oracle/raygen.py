"""Oracle: the step before the path -- patch pixel sampling, collation, perspective ray generation.
TEST INFRASTRUCTURE ONLY.  Pinned by tests/golden/raygen.npz (tests/golden/make_golden_raygen.py runs the reference's
PatchPixelSampler / RayGenerator / Cameras)."""
from typing import Dict, Tuple

import numpy as np
import torch


def patch_pixel_indices(u: torch.Tensor, num_images: int, height: int, width: int, patch: int) -> torch.Tensor:
    """PatchPixelSampler.sample_method without a mask, data/pixel_samplers.py:417-438.  u[P,3] -> int64 [P*patch^2, 3]."""
    sub = u.shape[0]
    idx = u * torch.tensor([num_images, height - patch, width - patch])
    idx = idx.view(sub, 1, 1, 3).broadcast_to(sub, patch, patch, 3).clone()
    yys, xxs = torch.meshgrid(torch.arange(patch), torch.arange(patch), indexing="ij")
    idx[:, ..., 1] += yys
    idx[:, ..., 2] += xxs
    return torch.floor(idx).long().flatten(0, 2)


def collate(images: torch.Tensor, image_idx: torch.Tensor, is_thermal_cameras: torch.Tensor,
            indices: torch.Tensor) -> Dict[str, torch.Tensor]:
    """collate_image_dataset_batch, data/pixel_samplers.py:239-256.  Reference quirk kept: on the CPU path `c` is a VIEW
    of `indices` (torch.split + flatten + a no-op .cpu()), so the in-place camera remap of :247 also changes the `c`
    used by the is_thermal lookup of :252-254, which therefore reads thermal_idx[image_idx[c]]."""
    c, y, x = indices[:, 0].clone(), indices[:, 1], indices[:, 2]
    out = {"image": images[c, y, x]}
    idx = indices.clone()
    idx[:, 0] = image_idx[c]
    out["indices"] = idx
    out["is_thermal"] = is_thermal_cameras[image_idx.sort()[1]][idx[:, 0]]
    return out


def generate_rays(indices: torch.Tensor, c2w: torch.Tensor, fx, fy, cx, cy) -> Tuple[torch.Tensor, ...]:
    """RayGenerator.forward + Cameras._generate_rays_from_coords for undistorted PERSPECTIVE cameras
    (model_components/ray_generators.py:40-55, cameras/cameras.py:600-905).  Returns origins, directions,
    pixel_area [R,1], directions_norm [R,1]."""
    c = indices[:, 0]
    y = indices[:, 1].float() + 0.5
    x = indices[:, 2].float() + 0.5
    fx, fy, cx, cy = fx[c], fy[c], cx[c], cy[c]
    coord = torch.stack([(x - cx) / fx, (y - cy) / fy], -1)
    coord_x = torch.stack([(x - cx + 1) / fx, (y - cy) / fy], -1)
    coord_y = torch.stack([(x - cx) / fx, (y - cy + 1) / fy], -1)
    cs = torch.stack([coord, coord_x, coord_y], dim=0)
    cs[..., 1] *= -1
    d = torch.empty((3, indices.shape[0], 3))
    d[..., 0] = cs[..., 0]
    d[..., 1] = cs[..., 1]
    d[..., 2] = -1.0
    m = c2w[c]
    d = torch.sum(d[..., None, :] * m[..., :3, :3], dim=-1)
    eps = torch.tensor([np.finfo(float).eps * 4.0]).to(d)
    norm = torch.maximum(torch.linalg.vector_norm(d, dim=-1, keepdims=True), eps)
    d = d / norm
    dx = torch.sqrt(torch.sum((d[0] - d[1]) ** 2, dim=-1))
    dy = torch.sqrt(torch.sum((d[0] - d[2]) ** 2, dim=-1))
    return m[..., :3, 3], d[0], (dx * dy)[..., None], norm[0]
