"""Per-ray losses of thermal-nerfacto (the row after the hot path).

Entry points with the reference's names: `interlevel_loss` / `distortion_loss`
(`nerfstudio/model_components/losses.py:114-158`) run as warp-per-ray kernels (csrc/tn_fused.cu: value and
gradient from one launch, no [R,S,S] temporaries); `tv_pixel_loss` / `cross_channel_loss` (`:593-651`) are the
torch fallbacks of the fused pixel-loss kernel (`fused_ops.pixel_losses`), used when the batch is not
patch-ordered.
"""
import torch
from torch import Tensor

L1Loss = torch.nn.L1Loss
MSELoss = torch.nn.MSELoss


def ray_samples_to_sdist(ray_samples) -> Tensor:
    """losses.py:106-111."""
    if getattr(ray_samples, "_layout", None) is not None:
        return getattr(ray_samples, "_layout").sbins
    starts, ends = ray_samples.spacing_starts, ray_samples.spacing_ends
    return torch.cat([starts[..., 0], ends[..., -1:, 0]], dim=-1)


def interlevel_loss(weights_list, ray_samples_list) -> Tensor:
    """losses.py:114-135: the final histogram (detached) must be bounded by every proposal histogram.  One
    warp-per-ray kernel per proposal level computes the value and the gradient w.r.t. the proposal weights."""
    from . import fused_ops

    c = ray_samples_to_sdist(ray_samples_list[-1]).detach()
    w = weights_list[-1][..., 0].detach()
    assert len(ray_samples_list) > 0
    loss_interlevel = 0.0
    for ray_samples, weights in zip(ray_samples_list[:-1], weights_list[:-1]):
        cp = ray_samples_to_sdist(ray_samples)
        wp = weights[..., 0]
        loss_interlevel = loss_interlevel + fused_ops.interlevel_loss_level(w, c, wp, cp)
    assert isinstance(loss_interlevel, Tensor)
    return loss_interlevel


def distortion_loss(weights_list, ray_samples_list) -> Tensor:
    """losses.py:153-158: mean over rays of lossfun_distortion on the final level.  The O(S^2) pairwise term runs
    in one warp-per-ray kernel (no [R,S,S] temporaries)."""
    from . import fused_ops

    c = ray_samples_to_sdist(ray_samples_list[-1])
    w = weights_list[-1][..., 0]
    return fused_ops.distortion_loss_rays(w, c)


def _rgb_patch_mask(is_thermal: Tensor) -> Tensor:
    """[R/4] float mask of the 2x2 patches that belong to RGB cameras.  The reference selects the RGB rays
    with a boolean index (a device->host sync) and views them as patches; with the patch-ordered batches
    its pixel sampler produces (data/pixel_samplers.py:389-438) a per-patch mask gives the same mean."""
    return (1 - is_thermal).view(-1, 4)[:, 0]


def _masked_mean(values: Tensor, mask: Tensor) -> Tensor:
    return (values * mask).sum() / mask.sum()


def tv_pixel_loss(pred_thermal: Tensor, is_thermal: Tensor) -> Tensor:
    """losses.py:603-620: total variation of the rendered thermal channel inside each 2x2 RGB-camera patch."""
    q = pred_thermal.reshape(-1, 4)
    tv = (q[:, 0] - q[:, 1]).abs() + (q[:, 0] - q[:, 2]).abs() + (q[:, 1] - q[:, 3]).abs() + (q[:, 2] - q[:, 3]).abs()
    return 0.25 * _masked_mean(tv, _rgb_patch_mask(is_thermal))


def pixel_grad(img: Tensor, patch_size: int = 2) -> Tensor:
    """losses.py:623-634: the four finite differences of a 2x2 patch, stacked [4, num_patches]."""
    assert patch_size == 2
    q = img.reshape(-1, 4)
    return torch.stack((q[:, 1] - q[:, 0], q[:, 2] - q[:, 0], q[:, 3] - q[:, 1], q[:, 3] - q[:, 2]))


def cross_channel_loss(pred_thermal: Tensor, gt_rgb: Tensor, is_thermal: Tensor) -> Tensor:
    """losses.py:637-651: L1 between patch gradients of rendered thermal and of grey-scale ground truth."""
    diff = (pixel_grad(pred_thermal) - pixel_grad(gt_rgb.mean(-1, keepdim=True))).abs().sum(0)
    return 0.25 * _masked_mean(diff, _rgb_patch_mask(is_thermal))
