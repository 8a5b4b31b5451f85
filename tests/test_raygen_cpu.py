"""SURVEY 8f-3 on CPU: the oracle's patch sampling / collation / ray generation against the reference fixture."""
import torch

import oracle.raygen as orag


def test_oracle_raygen_matches_reference(golden):
    g = golden("raygen.npz")
    H, W, patch = int(g["H"]), int(g["W"]), int(g["patch"])
    raw = orag.patch_pixel_indices(g["u"], g["images"].shape[0], H, W, patch)
    assert torch.equal(raw, g["raw_indices"])
    col = orag.collate(g["images"], g["image_idx"], g["is_thermal_cameras"], raw)
    assert torch.equal(col["indices"], g["indices"])
    assert torch.equal(col["image"], g["image"])
    assert torch.equal(col["is_thermal"], g["is_thermal"])
    o, d, area, norm = orag.generate_rays(g["indices"], g["c2w"], g["fx"], g["fy"], g["cx"], g["cy"])
    assert torch.equal(o, g["origins"])
    torch.testing.assert_close(d, g["directions"], rtol=0, atol=1e-7)
    torch.testing.assert_close(area, g["pixel_area"], rtol=1e-5, atol=0)
    torch.testing.assert_close(norm, g["directions_norm"], rtol=1e-6, atol=0)
    # patches: groups of patch^2 consecutive rays share the camera and form a patch x patch block
    idx = g["indices"].view(-1, patch * patch, 3)
    assert (idx[:, :, 0] == idx[:, :1, 0]).all()
    assert (idx[:, :, 1].max(1).values - idx[:, :, 1].min(1).values == patch - 1).all()


def test_raygen_module_guards():
    import pytest

    from nerfstudio_thermal_b200 import raygen
    c2w = torch.eye(4)[None, :3, :4]
    with pytest.raises(NotImplementedError):
        raygen.Cameras(c2w, 1.0, 1.0, 0.5, 0.5, 4, 4, camera_type="FISHEYE")
    with pytest.raises(NotImplementedError):
        raygen.Cameras(c2w, 1.0, 1.0, 0.5, 0.5, 4, 4, distortion_params=torch.tensor([[0.1, 0, 0, 0, 0, 0]]))
    cams = raygen.Cameras(c2w.repeat(3, 1, 1), torch.tensor([1.0, 2.0, 3.0]), 2.0, 0.5, 0.5, 4, 4)
    assert cams.intrinsics.shape == (3, 4) and len(cams) == 3
    s = raygen.PatchPixelSampler(patch_size=2, num_rays_per_batch=4099)
    assert s.num_rays_per_batch == 4096
    with pytest.raises(RuntimeError, match="GPU"):
        s.sample({"image": torch.zeros(2, 8, 8, 3), "image_idx": torch.arange(2)})
