"""Oracle: volume-rendering reductions.  TEST INFRASTRUCTURE ONLY."""
from typing import Union

import torch

_NAMED = {"white": 1.0, "black": 0.0}


def render_colour(colour: torch.Tensor, weights: torch.Tensor, background: Union[str, torch.Tensor] = "last_sample",
                  training: bool = True) -> torch.Tensor:
    """RGBRenderer / RGBTRenderer forward + combine_rgb (dense samples),
    model_components/renderers.py:118-133, 238-245 (and :292-307, 418-425 for RGBT).

    colour [R,S,C], weights [R,S,1] -> [R,C].
    """
    if not training:
        colour = torch.nan_to_num(colour)
    comp = torch.sum(weights * colour, dim=-2)
    acc = torch.sum(weights, dim=-2)
    if isinstance(background, str) and background == "random":
        out = comp
    else:
        if isinstance(background, str) and background == "last_sample":
            bg = colour[..., -1, :]
        elif isinstance(background, str):
            bg = torch.full_like(comp, _NAMED[background])
            if comp.shape[-1] == 4:  # utils/colors.py:37-48: named RGBT colours have T = 0
                bg[..., 3] = 0.0
        else:
            bg = background.expand(comp.shape)
        out = comp + bg * (1.0 - acc)
    if not training:
        out = torch.clamp(out, min=0.0, max=1.0)
    return out


def render_accumulation(weights: torch.Tensor) -> torch.Tensor:
    """AccumulationRenderer, renderers.py:509."""
    return torch.sum(weights, dim=-2)


def render_depth_median(weights: torch.Tensor, starts: torch.Tensor, ends: torch.Tensor) -> torch.Tensor:
    """DepthRenderer(method="median"), renderers.py:547-557."""
    steps = (starts + ends) / 2
    cum = torch.cumsum(weights[..., 0], dim=-1)
    split = torch.ones((*weights.shape[:-2], 1), device=weights.device) * 0.5
    idx = torch.searchsorted(cum, split, side="left")
    idx = torch.clamp(idx, 0, steps.shape[-2] - 1)
    return torch.gather(steps[..., 0], dim=-1, index=idx)


def render_depth_expected(weights: torch.Tensor, starts: torch.Tensor, ends: torch.Tensor) -> torch.Tensor:
    """DepthRenderer(method="expected"), renderers.py:558-576 (batch-global clip)."""
    steps = (starts + ends) / 2
    depth = torch.sum(weights * steps, dim=-2) / (torch.sum(weights, -2) + 1e-10)
    return torch.clip(depth, steps.min(), steps.max())
