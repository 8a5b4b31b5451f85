"""The step before the path, on the device: patch pixel sampling, collation and ray generation.

Mirrors `PatchPixelSampler` (data/pixel_samplers.py:360-438), the collation of `PixelSampler.collate_image_dataset_batch`
(:225-256) and `RayGenerator` over `Cameras` (model_components/ray_generators.py:24-55, cameras/cameras.py:504-905) for
undistorted perspective cameras -- the thermal-nerfacto data path.  In the reference the sampled (c, y, x) indices make
a device -> CPU round trip every step to index CPU image tensors (:239-240) and the rays are built by ~60 torch ops;
here the cached images live in HBM and three launches produce the whole train batch, so a step needs no host input
at all.  Other camera models / distortion parameters raise at construction.
"""
from dataclasses import dataclass
from typing import Dict, Optional

import torch
from torch import Tensor

from ._lib import call, ptr, stream
from .rays import RayBundle


@dataclass
class Cameras:
    """The fields of cameras/cameras.py:Cameras that perspective ray generation reads."""

    camera_to_worlds: Tensor  # [C,3,4]
    fx: Tensor
    fy: Tensor
    cx: Tensor
    cy: Tensor
    width: int
    height: int
    distortion_params: Optional[Tensor] = None
    camera_type: str = "PERSPECTIVE"

    def __post_init__(self):
        if self.camera_type != "PERSPECTIVE":
            raise NotImplementedError(f"camera type {self.camera_type}: only PERSPECTIVE rays are generated on the device")
        if self.distortion_params is not None and bool((self.distortion_params != 0).any()):
            raise NotImplementedError("non-zero distortion parameters: undistortion is not part of this path")
        n = self.camera_to_worlds.shape[0]
        full = lambda t: torch.as_tensor(t, dtype=torch.float32).reshape(-1).expand(n)  # noqa: E731
        self.intrinsics = torch.stack([full(self.fx), full(self.fy), full(self.cx), full(self.cy)], dim=-1).contiguous()

    def __len__(self) -> int:
        return self.camera_to_worlds.shape[0]

    def to(self, device) -> "Cameras":
        out = Cameras(self.camera_to_worlds.to(device), self.fx, self.fy, self.cx, self.cy, self.width, self.height,
                      None, self.camera_type)
        out.intrinsics = self.intrinsics.to(device)
        return out


class RayGenerator(torch.nn.Module):
    """model_components/ray_generators.py:24-55: ray_indices [R,3] (camera, row, col) -> RayBundle, one launch."""

    def __init__(self, cameras: Cameras) -> None:
        super().__init__()
        self.cameras = cameras
        self.register_buffer("c2w", cameras.camera_to_worlds.float().contiguous(), persistent=False)
        self.register_buffer("intrinsics", cameras.intrinsics.float().contiguous(), persistent=False)

    def forward(self, ray_indices: Tensor) -> RayBundle:
        idx = ray_indices.to(torch.int64).contiguous()
        r = idx.shape[0]
        dev = self.c2w.device
        origins = torch.empty((r, 3), device=dev)
        directions = torch.empty((r, 3), device=dev)
        area = torch.empty((r, 1), device=dev)
        norm = torch.empty((r, 1), device=dev)
        cams = torch.empty((r, 1), device=dev, dtype=torch.int64)
        call("tn_generate_rays", ptr(idx), ptr(self.c2w), ptr(self.intrinsics), self.c2w.shape[0], r, ptr(origins),
             ptr(directions), ptr(area), ptr(norm), ptr(cams), stream())
        return RayBundle(origins=origins, directions=directions, pixel_area=area, camera_indices=cams,
                         metadata={"directions_norm": norm})


class PatchPixelSampler:
    """PatchPixelSampler (data/pixel_samplers.py:360-438) over a cached image batch that lives on the device.

    image_batch: {"image": [n,H,W,3] float32 or uint8, "image_idx": [n] camera index of every cached image,
    "is_thermal": [num_cameras] flags}.  sample() returns {"image" [R,3], "indices" [R,3] (camera, row, col),
    "is_thermal" [R]} -- two launches and one `torch.rand`, nothing leaves the device."""

    def __init__(self, patch_size: int = 2, num_rays_per_batch: int = 4096) -> None:
        self.patch_size = patch_size
        self.set_num_rays_per_batch(num_rays_per_batch)

    def set_num_rays_per_batch(self, num_rays_per_batch: int) -> None:
        """pixel_samplers.py:379-386: a whole number of patches."""
        self.num_rays_per_batch = (num_rays_per_batch // (self.patch_size**2)) * (self.patch_size**2)

    def sample_method(self, batch_size: int, num_images: int, image_height: int, image_width: int, device,
                      u: Optional[Tensor] = None) -> Tensor:
        sub = batch_size // (self.patch_size**2)
        if u is None:
            u = torch.rand((sub, 3), device=device)
        u = u.to(device=device, dtype=torch.float32).contiguous()
        out = torch.empty((sub * self.patch_size**2, 3), device=device, dtype=torch.int64)
        call("tn_patch_pixel_indices", ptr(u), sub, self.patch_size, num_images, image_height, image_width, ptr(out),
             stream())
        return out

    def sample(self, image_batch: Dict[str, Tensor], u: Optional[Tensor] = None) -> Dict[str, Tensor]:
        images = image_batch["image"]
        if not images.is_cuda:
            raise RuntimeError("the cached image batch must live on the GPU (no CPU fallback)")
        if images.dtype not in (torch.float32, torch.uint8):
            raise TypeError(f"image dtype {images.dtype}: float32 or uint8")
        images = images.contiguous()
        n, h, w, ch = images.shape
        dev = images.device
        image_idx = image_batch["image_idx"].to(device=dev, dtype=torch.int64).contiguous()
        indices = self.sample_method(self.num_rays_per_batch, n, h, w, dev, u)
        r = indices.shape[0]
        thermal_tbl = None
        if "is_thermal" in image_batch:
            # pixel_samplers.py:252-254 looks the flag up as is_thermal[argsort(image_idx)][c'], where c' is the
            # ALREADY REMAPPED camera index image_idx[c] (on the reference's CPU path `c` aliases `indices`, which
            # :247 rewrites in place); the per-image table below folds both lookups
            flags = image_batch["is_thermal"].to(device=dev, dtype=torch.float32)
            thermal_tbl = flags[torch.argsort(image_idx, stable=True)][image_idx].contiguous()
        image = torch.empty((r, ch), device=dev)
        is_thermal = torch.empty((r,), device=dev) if thermal_tbl is not None else None
        call("tn_gather_pixels", ptr(images), 0 if images.dtype == torch.float32 else 1, n, h, w, ch, ptr(indices),
             ptr(image_idx), ptr(thermal_tbl), r, ptr(image), ptr(is_thermal), stream())
        out = {"image": image, "indices": indices}
        if is_thermal is not None:
            out["is_thermal"] = is_thermal
        return out
