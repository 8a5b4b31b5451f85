"""Renderers with the reference's class names and call signatures.

Mirrors `nerfstudio/model_components/renderers.py`: RGBRenderer (:74-245), RGBTRenderer (:248-425),
AccumulationRenderer (:482-510), DepthRenderer (:513-578), for dense [R,S,*] samples (the packed
`ray_indices` path belongs to nerfacc-based models and is rejected).
"""
from typing import Literal, Optional, Tuple, Union

import torch
from torch import Tensor, nn

from . import ops
from .rays import RaySamples

BackgroundColor = Union[Literal["random", "last_sample", "black", "white"], Tensor]

# utils/colors.py:22-48
COLORS_DICT = {"white": (1.0, 1.0, 1.0), "black": (0.0, 0.0, 0.0), "red": (1.0, 0.0, 0.0), "green": (0.0, 1.0, 0.0),
               "blue": (0.0, 0.0, 1.0)}
COLORS_RGBT_DICT = {k: v + (0.0,) for k, v in COLORS_DICT.items()}


def rgb_to_rgbt_image(image: Tensor, is_thermal: Tensor) -> Tensor:
    """utils/rgbt_utils.py:6-33."""
    rgbt = torch.zeros(image.shape[:-1] + (4,), device=image.device)
    is_rgb = 1 - is_thermal
    if hasattr(is_rgb, "__len__"):
        rgbt[..., :3] = torch.einsum("ij,i->ij", image, is_rgb)
    else:
        rgbt[..., :3] += image * is_rgb
    rgbt[..., 3] = image[..., 0] * is_thermal
    return rgbt


def _no_packed(ray_indices, num_rays):
    if ray_indices is not None or num_rays is not None:
        raise NotImplementedError("packed samples (ray_indices/num_rays) are not part of the thermal-nerfacto path")


class RGBRenderer(nn.Module):
    """renderers.py:74-245."""

    _palette = COLORS_DICT

    def __init__(self, background_color: BackgroundColor = "random", num_channels: int = 3) -> None:
        super().__init__()
        self.background_color: BackgroundColor = background_color
        self.num_channels = num_channels

    def _bg_args(self, background_color, channels: int):
        if isinstance(background_color, str):
            if background_color == "random":
                return ops.BG_NONE, None, None
            if background_color == "last_sample":
                return ops.BG_LAST_SAMPLE, None, None
            if background_color in self._palette:
                col = self._palette[background_color]
                assert len(col) == channels or channels <= len(col), "Background color must match the channel count"
                return ops.BG_CONSTANT, tuple(col[:channels]), None
            raise ValueError(f"unknown background colour {background_color!r}")
        assert isinstance(background_color, Tensor)
        if background_color.numel() == channels:
            return ops.BG_CONSTANT, tuple(float(v) for v in background_color.reshape(-1).tolist()), None
        return ops.BG_NONE, None, background_color  # per-ray background: blended after the kernel

    def combine_rgb(self, rgb: Tensor, weights: Tensor, background_color: BackgroundColor = "random",
                    ray_indices: Optional[Tensor] = None, num_rays: Optional[int] = None,
                    eval_mode: bool = False) -> Tensor:
        """renderers.py:86-133: sum_i w_i rgb_i (+ background * (1 - sum_i w_i))."""
        _no_packed(ray_indices, num_rays)
        lead, s, c = rgb.shape[:-2], rgb.shape[-2], rgb.shape[-1]
        mode, const, per_ray = self._bg_args(background_color, c)
        comp, acc, _, _, _ = ops.render(weights.reshape(-1, s), rgb.reshape(-1, s, c), None, None, bg_mode=mode,
                                        bg=const, eval_mode=eval_mode and per_ray is None)
        comp = comp.view(*lead, c)
        if per_ray is not None:
            comp = comp + per_ray.expand(comp.shape).to(comp.device) * (1.0 - acc.view(*lead, 1))
            if eval_mode:
                comp = torch.clamp(comp, min=0.0, max=1.0)
        return comp

    def get_background_color(self, background_color: BackgroundColor, shape: Tuple[int, ...], device) -> Tensor:
        assert background_color not in {"last_sample", "random"}
        assert shape[-1] == self.num_channels, "Background color must be RGB."
        if isinstance(background_color, str) and background_color in self._palette:
            background_color = torch.tensor(self._palette[background_color])
        assert isinstance(background_color, Tensor)
        return background_color.expand(shape).to(device)

    def blend_background(self, image: Tensor, background_color: Optional[BackgroundColor] = None) -> Tensor:
        """renderers.py:161-187."""
        if image.size(-1) < self.num_channels + 1:
            return image
        rgb, opacity = image[..., : self.num_channels], image[..., self.num_channels:]
        if background_color is None:
            background_color = self.background_color
            if background_color in {"last_sample", "random"}:
                background_color = "black"
        background_color = self.get_background_color(background_color, shape=rgb.shape, device=rgb.device)
        return rgb * opacity + background_color.to(rgb.device) * (1 - opacity)

    def blend_background_for_loss_computation(self, pred_image: Tensor, pred_accumulation: Tensor,
                                              gt_image: Tensor) -> Tuple[Tensor, Tensor]:
        """renderers.py:189-212."""
        background_color = self.background_color
        if background_color == "last_sample":
            background_color = "black"
        elif background_color == "random":
            background_color = torch.rand_like(pred_image)
            pred_image = pred_image + background_color * (1.0 - pred_accumulation)
        gt_image = self.blend_background(gt_image, background_color=background_color)
        return pred_image, gt_image

    def forward(self, rgb: Tensor, weights: Tensor, ray_indices: Optional[Tensor] = None,
                num_rays: Optional[int] = None, background_color: Optional[BackgroundColor] = None) -> Tensor:
        """renderers.py:214-245 (eval: nan_to_num on the samples, clamp of the composite -- both in-kernel)."""
        if background_color is None:
            background_color = self.background_color
        return self.combine_rgb(rgb, weights, background_color=background_color, ray_indices=ray_indices,
                                num_rays=num_rays, eval_mode=not self.training)


class RGBTRenderer(RGBRenderer):
    """renderers.py:248-425: four channels, named background colours have T = 0."""

    _palette = COLORS_RGBT_DICT

    def __init__(self, background_color: BackgroundColor = "random") -> None:
        super().__init__(background_color=background_color, num_channels=4)

    def blend_background(self, image: Tensor, is_thermal: Tensor,  # type: ignore[override]
                         background_color: Optional[BackgroundColor] = None) -> Tensor:
        """renderers.py:336-365."""
        if image.size(-1) < 4:
            return rgb_to_rgbt_image(image, is_thermal)
        opacity = image[..., 3:]
        rgbt = rgb_to_rgbt_image(image, is_thermal)
        if background_color is None:
            background_color = self.background_color
            if background_color in {"last_sample", "random"}:
                background_color = "black"
        background_color = self.get_background_color(background_color, shape=rgbt.shape, device=rgbt.device)
        return rgbt * opacity + background_color.to(rgbt.device) * (1 - opacity)

    def blend_background_for_loss_computation(self, pred_image: Tensor, pred_accumulation: Tensor,  # type: ignore[override]
                                              gt_image: Tensor, is_thermal: Tensor) -> Tuple[Tensor, Tensor]:
        """renderers.py:367-392."""
        background_color = self.background_color
        if background_color == "last_sample":
            background_color = "black"
        elif background_color == "random":
            background_color = torch.rand_like(pred_image)
            pred_image = pred_image + background_color * (1.0 - pred_accumulation)
        gt_image = self.blend_background(gt_image, is_thermal, background_color=background_color)
        return pred_image, gt_image


class AccumulationRenderer(nn.Module):
    """renderers.py:482-510."""

    @classmethod
    def forward(cls, weights: Tensor, ray_indices: Optional[Tensor] = None, num_rays: Optional[int] = None) -> Tensor:
        _no_packed(ray_indices, num_rays)
        lead, s = weights.shape[:-2], weights.shape[-2]
        _, acc, _, _, _ = ops.render(weights.reshape(-1, s), None, None, None)
        return acc.view(*lead, 1)


class DepthRenderer(nn.Module):
    """renderers.py:513-578.  "expected" keeps the reference's batch-global clip to [steps.min(), steps.max()]
    (:574): the kernel returns the extrema of the launch, the clip is a tensor-tensor clamp (no host sync)."""

    def __init__(self, method: Literal["median", "expected"] = "median") -> None:
        super().__init__()
        self.method = method

    def forward(self, weights: Tensor, ray_samples: RaySamples, ray_indices: Optional[Tensor] = None,
                num_rays: Optional[int] = None) -> Tensor:
        _no_packed(ray_indices, num_rays)
        if self.method not in ("median", "expected"):
            raise NotImplementedError(f"Method {self.method} not implemented")
        lead, s = weights.shape[:-2], weights.shape[-2]
        lay = getattr(ray_samples, "_layout", None)
        if lay is not None and lay.ebins.shape[-1] == s + 1 and len(lead) == 1:
            starts = ends = None  # sampler-made samples: the kernel reads the bin edges in place
            bins = lay.ebins
        else:
            starts = ray_samples.frustums.starts.reshape(-1, s)
            ends = ray_samples.frustums.ends.reshape(-1, s)
            bins = None
        w = weights.reshape(-1, s)
        if self.method == "median":
            _, _, med, _, _ = ops.render(w.detach(), None, starts, ends, want_depth=True, bins=bins)
            return med.view(*lead, 1)
        _, _, _, exp, minmax = ops.render(w, None, starts, ends, want_depth=True, bins=bins)
        return torch.clamp(exp, minmax[0], minmax[1]).view(*lead, 1)
