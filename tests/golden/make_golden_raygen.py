"""Golden fixture for the step before the path: tests/golden/raygen.npz.

Runs the UNMODIFIED reference: `PatchPixelSampler.sample_method` + the collation of `collate_image_dataset_batch`
(data/pixel_samplers.py:239-256, 417-438) and `RayGenerator` over `Cameras` (model_components/ray_generators.py,
cameras/cameras.py) for undistorted perspective cameras.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_raygen.py
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_shim  # noqa: E402

_ref_shim.install()

from nerfstudio.cameras.cameras import Cameras, CameraType  # noqa: E402
from nerfstudio.data.pixel_samplers import PatchPixelSamplerConfig  # noqa: E402
from nerfstudio.model_components.ray_generators import RayGenerator  # noqa: E402


def random_rotation(gen):
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=gen))
    return q * torch.sign(torch.linalg.det(q))


def main():
    gen = torch.Generator().manual_seed(11)
    n_cam, H, W, patch, rays = 6, 48, 64, 2, 256
    c2w = torch.zeros(n_cam, 3, 4)
    for i in range(n_cam):
        c2w[i, :, :3] = random_rotation(gen)
        c2w[i, :, 3] = torch.randn(3, generator=gen)
    fx = 50.0 + 10 * torch.rand(n_cam, generator=gen)
    fy = 52.0 + 10 * torch.rand(n_cam, generator=gen)
    cx = W / 2 + torch.randn(n_cam, generator=gen)
    cy = H / 2 + torch.randn(n_cam, generator=gen)
    cams = Cameras(camera_to_worlds=c2w, fx=fx, fy=fy, cx=cx, cy=cy, width=W, height=H,
                   camera_type=CameraType.PERSPECTIVE)
    images = torch.rand(n_cam, H, W, 3, generator=gen)
    image_idx = torch.tensor([3, 0, 5, 1, 4, 2])           # the dataloader's shuffled image order
    is_thermal = torch.tensor([0.0, 0.0, 0.0, 1.0, 1.0, 1.0])  # per CAMERA index
    sampler = PatchPixelSamplerConfig(patch_size=patch, num_rays_per_batch=rays).setup()
    torch.manual_seed(5)
    u = torch.rand((rays // patch**2, 3))                   # what sample_method is about to draw
    torch.manual_seed(5)
    batch = sampler.collate_image_dataset_batch({"image": images, "image_idx": image_idx, "is_thermal": is_thermal},
                                                rays)
    torch.manual_seed(5)
    raw = sampler.sample_method(rays, n_cam, H, W)          # (position in batch, y, x) before the camera remap
    bundle = RayGenerator(cams)(batch["indices"])
    out = {
        "c2w": c2w, "fx": fx, "fy": fy, "cx": cx, "cy": cy, "H": np.array(H), "W": np.array(W), "patch": np.array(patch),
        "images": images, "image_idx": image_idx, "is_thermal_cameras": is_thermal, "u": u, "raw_indices": raw,
        "indices": batch["indices"], "image": batch["image"], "is_thermal": batch["is_thermal"],
        "origins": bundle.origins, "directions": bundle.directions, "pixel_area": bundle.pixel_area,
        "camera_indices": bundle.camera_indices, "directions_norm": bundle.metadata["directions_norm"],
    }
    path = os.path.join(HERE, "raygen.npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else v) for k, v in out.items()})
    print(f"wrote raygen.npz: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
