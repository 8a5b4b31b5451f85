"""Oracle: ray sampling, PDF resampling and sample weights.  TEST INFRASTRUCTURE ONLY."""
from dataclasses import dataclass
from typing import Optional

import torch


def piecewise_spacing(x: torch.Tensor) -> torch.Tensor:
    """UniformLinDispPiecewiseSampler spacing_fn, model_components/ray_samplers.py:244."""
    return torch.where(x < 1, x / 2, 1 - 1 / (2 * x))


def piecewise_spacing_inv(x: torch.Tensor) -> torch.Tensor:
    """ray_samplers.py:245."""
    return torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x))


@dataclass
class OracleSamples:
    """The subset of RaySamples/Frustums (cameras/rays.py:32-150) the path reads."""

    origins: torch.Tensor  # [R,3]
    directions: torch.Tensor  # [R,3]
    camera_indices: Optional[torch.Tensor]  # [R,1] int64
    starts: torch.Tensor  # [R,S,1] euclidean bin starts
    ends: torch.Tensor  # [R,S,1]
    spacing_starts: torch.Tensor  # [R,S,1] normalised bins in [0,1]
    spacing_ends: torch.Tensor  # [R,S,1]
    s_near: torch.Tensor  # [R,1]
    s_far: torch.Tensor  # [R,1]

    @property
    def deltas(self) -> torch.Tensor:  # rays.py:268
        return self.ends - self.starts

    def spacing_to_euclidean(self, x: torch.Tensor) -> torch.Tensor:  # ray_samplers.py:115-116
        return piecewise_spacing_inv(x * self.s_far + (1 - x) * self.s_near)

    def sdist(self) -> torch.Tensor:  # model_components/losses.py:106-111
        return torch.cat([self.spacing_starts[..., 0], self.spacing_ends[..., -1:, 0]], dim=-1)


def _make_samples(origins, directions, camera_indices, bins, s_near, s_far) -> OracleSamples:
    """RayBundle.get_ray_samples, cameras/rays.py:251-295."""
    euclid = piecewise_spacing_inv(bins * s_far + (1 - bins) * s_near)
    if bins.shape[0] != origins.shape[0]:
        bins = bins.expand(origins.shape[0], -1)
    return OracleSamples(
        origins=origins, directions=directions, camera_indices=camera_indices,
        starts=euclid[..., :-1, None], ends=euclid[..., 1:, None],
        spacing_starts=bins[..., :-1, None], spacing_ends=bins[..., 1:, None],
        s_near=s_near, s_far=s_far,
    )


def initial_bins(num_rays: int, num_samples: int, jitter: Optional[torch.Tensor], device=None) -> torch.Tensor:
    """SpacedSampler.generate_ray_samples, ray_samplers.py:100-111.

    jitter: None (eval) or the [R,1] tensor the reference draws with torch.rand (single_jitter=True).
    """
    bins = torch.linspace(0.0, 1.0, num_samples + 1).to(device)[None, ...]  # made on the CPU, then moved (:100)
    if jitter is not None:
        centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
        upper = torch.cat([centers, bins[..., -1:]], -1)
        lower = torch.cat([bins[..., :1], centers], -1)
        bins = lower + (upper - lower) * jitter
    return bins


def initial_samples(origins, directions, camera_indices, nears, fars, num_samples, jitter) -> OracleSamples:
    bins = initial_bins(origins.shape[0], num_samples, jitter, origins.device)
    return _make_samples(origins, directions, camera_indices, bins, piecewise_spacing(nears), piecewise_spacing(fars))


def pdf_resample(prev: OracleSamples, weights: torch.Tensor, num_samples: int, jitter: Optional[torch.Tensor],
                 histogram_padding: float = 0.01, eps: float = 1e-5) -> OracleSamples:
    """PDFSampler.generate_ray_samples (include_original=False), ray_samplers.py:301-372.

    weights [R,S_old,1]; jitter None (eval: bin centres) or [R,1] torch.rand draw (train, single jitter).
    """
    num_bins = num_samples + 1
    w = weights[..., 0] + histogram_padding
    w_sum = torch.sum(w, dim=-1, keepdim=True)
    padding = torch.relu(eps - w_sum)
    w = w + padding / w.shape[-1]
    w_sum = w_sum + padding
    pdf = w / w_sum
    cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, dim=-1))
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)

    u = torch.linspace(0.0, 1.0 - (1.0 / num_bins), steps=num_bins).to(cdf.device)
    if jitter is not None:
        u = u.expand(size=(*cdf.shape[:-1], num_bins))
        u = u + jitter / num_bins
    else:
        u = u + 1.0 / (2 * num_bins)
        u = u.expand(size=(*cdf.shape[:-1], num_bins))
    u = u.contiguous()

    existing = prev.sdist()
    inds = torch.searchsorted(cdf, u, side="right")
    below = torch.clamp(inds - 1, 0, existing.shape[-1] - 1)
    above = torch.clamp(inds, 0, existing.shape[-1] - 1)
    cdf0 = torch.gather(cdf, -1, below)
    bin0 = torch.gather(existing, -1, below)
    cdf1 = torch.gather(cdf, -1, above)
    bin1 = torch.gather(existing, -1, above)
    t = torch.clip(torch.nan_to_num((u - cdf0) / (cdf1 - cdf0), 0), 0, 1)
    bins = (bin0 + t * (bin1 - bin0)).detach()
    return _make_samples(prev.origins, prev.directions, prev.camera_indices, bins, prev.s_near, prev.s_far)


# Checker-only switch (conditioning experiments of the parity tests): move every sample position this many float32
# ulps towards +inf before it is used.  0 = the reference's arithmetic.
POSITION_ULP_SHIFT = 0


def sample_positions(s: OracleSamples) -> torch.Tensor:
    """Frustums.get_positions, cameras/rays.py:49-58 -> [R,S,3]."""
    pos = s.origins[:, None, :] + s.directions[:, None, :] * (s.starts + s.ends) / 2
    if POSITION_ULP_SHIFT:
        shifted = pos.detach()
        for _ in range(POSITION_ULP_SHIFT):
            shifted = torch.nextafter(shifted, torch.full_like(shifted, float("inf")))
        pos = pos + (shifted - pos.detach())  # same gradient path, values moved by whole ulps
    return pos


def sample_weights(deltas: torch.Tensor, densities: torch.Tensor) -> torch.Tensor:
    """RaySamples.get_weights, cameras/rays.py:128-150."""
    dd = deltas * densities
    alphas = 1 - torch.exp(-dd)
    trans = torch.cumsum(dd[..., :-1, :], dim=-2)
    trans = torch.cat([torch.zeros((*trans.shape[:1], 1, 1), device=densities.device), trans], dim=-2)
    trans = torch.exp(-trans)
    return torch.nan_to_num(alphas * trans)
