"""SURVEY 8f-3 kernels through the C ABI against the reference fixture: indices bit-exact, rays to 1 ulp."""
import pytest
import torch

import oracle.raygen as orag

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(g):
    from nerfstudio_thermal_b200 import raygen
    cams = raygen.Cameras(g["c2w"], g["fx"], g["fy"], g["cx"], g["cy"], int(g["W"]), int(g["H"])).to(DEV)
    return raygen, cams


@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8])
def test_patch_sampler_and_collation_match_reference(golden, dtype):
    g = golden("raygen.npz")
    raygen, _ = _setup(g)
    patch = int(g["patch"])
    images = g["images"] if dtype == torch.float32 else (g["images"] * 255).round().to(torch.uint8)
    sampler = raygen.PatchPixelSampler(patch_size=patch, num_rays_per_batch=g["indices"].shape[0])
    raw = sampler.sample_method(sampler.num_rays_per_batch, images.shape[0], int(g["H"]), int(g["W"]), DEV, u=g["u"])
    assert torch.equal(raw.cpu(), g["raw_indices"])
    batch = sampler.sample({"image": images.to(DEV), "image_idx": g["image_idx"].to(DEV),
                            "is_thermal": g["is_thermal_cameras"].to(DEV)}, u=g["u"])
    assert torch.equal(batch["indices"].cpu(), g["indices"])
    assert torch.equal(batch["is_thermal"].cpu(), g["is_thermal"])
    if dtype == torch.float32:
        assert torch.equal(batch["image"].cpu(), g["image"])
    else:
        torch.testing.assert_close(batch["image"].cpu(), g["image"], atol=0.5 / 255 + 1e-6, rtol=0)


def test_generate_rays_matches_reference(golden):
    g = golden("raygen.npz")
    raygen, cams = _setup(g)
    b = raygen.RayGenerator(cams).to(DEV)(g["indices"].to(DEV))
    assert torch.equal(b.origins.cpu(), g["origins"])
    assert torch.equal(b.camera_indices.cpu(), g["camera_indices"])
    torch.testing.assert_close(b.directions.cpu(), g["directions"], rtol=0, atol=1.2e-7)
    torch.testing.assert_close(b.pixel_area.cpu(), g["pixel_area"], rtol=2e-4, atol=0)
    torch.testing.assert_close(b.metadata["directions_norm"].cpu(), g["directions_norm"], rtol=1e-6, atol=0)


def test_device_batch_feeds_a_train_step_at_full_size():
    """4096 rays from 64 cached 640x512 images: sampler + ray generator -> batch -> properties, and the oracle on a slice."""
    from nerfstudio_thermal_b200 import raygen
    torch.manual_seed(2)
    n, H, W = 64, 512, 640
    c2w = torch.zeros(n, 3, 4)
    c2w[:, :, :3] = torch.linalg.qr(torch.randn(n, 3, 3))[0]
    c2w[:, :, 3] = torch.randn(n, 3) * 0.3
    fx = torch.full((n,), 520.0)
    cams = raygen.Cameras(c2w, fx, fx * 1.01, W / 2.0, H / 2.0, W, H)
    images = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device=DEV)
    sampler = raygen.PatchPixelSampler(2, 4096)
    batch = sampler.sample({"image": images, "image_idx": torch.arange(n, device=DEV),
                            "is_thermal": (torch.arange(n) >= 32).float().to(DEV)})
    idx = batch["indices"]
    assert idx.shape == (4096, 3) and int(idx[:, 0].max()) < n and int(idx[:, 1].max()) < H and int(idx[:, 2].max()) < W
    assert torch.equal(batch["is_thermal"], (idx[:, 0] >= 32).float())
    # true division as on the reference's CPU data path (torch's CUDA scalar division multiplies by a reciprocal)
    assert torch.equal(batch["image"].cpu(), images[idx[:, 0], idx[:, 1], idx[:, 2]].cpu().float() / 255.0)
    b = raygen.RayGenerator(cams.to(DEV)).to(DEV)(idx)
    torch.testing.assert_close(b.directions.norm(dim=-1), torch.ones(4096, device=DEV), rtol=0, atol=1e-6)
    sl = slice(0, 4096, 97)
    o, d, area, _ = orag.generate_rays(idx[sl].cpu(), c2w, cams.intrinsics[:, 0], cams.intrinsics[:, 1],
                                       cams.intrinsics[:, 2], cams.intrinsics[:, 3])
    assert torch.equal(b.origins[sl].cpu(), o)
    torch.testing.assert_close(b.directions[sl].cpu(), d, rtol=0, atol=1.2e-7)
    torch.testing.assert_close(b.pixel_area[sl].cpu(), area, rtol=2e-4, atol=0)


def test_graphed_train_step_from_device_source(golden):
    """The whole iteration -- patch sampling, collation, ray generation, forward, losses, backward, Adam -- replayed as
    one CUDA graph with no host input; every replay draws a new batch (the loss changes) and the loss trends down."""
    from nerfstudio_thermal_b200 import engine, optim, raygen
    from test_gpu_model import build
    g, model = build(golden, "separate")
    torch.manual_seed(4)
    n, H, W = 8, 32, 40
    c2w = torch.zeros(n, 3, 4)
    c2w[:, :, :3] = torch.linalg.qr(torch.randn(n, 3, 3))[0]
    c2w[:, :, 3] = torch.randn(n, 3) * 0.2
    cams = raygen.Cameras(c2w, 40.0, 40.0, W / 2.0, H / 2.0, W, H).to(DEV)
    images = torch.rand(n, H, W, 3, device=DEV)
    src = engine.DeviceBatchSource(raygen.PatchPixelSampler(2, 64), raygen.RayGenerator(cams).to(DEV),
                                   {"image": images, "image_idx": torch.arange(n, device=DEV),
                                    "is_thermal": torch.tensor([0.0] * 4 + [1.0] * 4, device=DEV)})
    example = src.next()
    runner = engine.GraphedTrainStep(model, example, source=src, optimizer=optim.thermal_nerfacto_optimizers())
    losses = [float(runner.step().detach()) for _ in range(12)]
    assert len(set(losses)) == len(losses)  # a fresh batch every replay
    assert all(torch.isfinite(torch.tensor(losses)))
    assert sum(losses[-4:]) < sum(losses[:4])
    assert runner.optimizer.step_count == 12
