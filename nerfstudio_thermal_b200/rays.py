"""Ray data structures mirroring the reference's `nerfstudio/cameras/rays.py` API.

Same attribute names, shapes and methods (`Frustums.get_positions`, `RaySamples.get_weights`,
`RayBundle.get_ray_samples`, cameras/rays.py:32-295), so code written against the reference reads the
same here.  The samplers of this package additionally attach a `RayLayout` (private `_layout`) to the
`RaySamples` they produce: the per-ray tensors ([R,3] origins/directions, [R,S+1] bin edges) from which
the public per-sample views were made.  Kernels use it to skip the materialised [R,S,3] tensors; a
user-built `RaySamples` without it takes the generic path.
"""
from dataclasses import dataclass, field, fields, replace
from typing import Callable, Dict, Optional

import torch
from torch import Tensor

from . import ops


@dataclass
class RayLayout:
    origins: Tensor  # [R,3]
    directions: Tensor  # [R,3]
    ebins: Tensor  # [R,S+1] euclidean bin edges
    sbins: Tensor  # [R,S+1] spacing-domain bin edges
    nears: Tensor  # [R]
    fars: Tensor  # [R]
    # Set by the model for the layouts of ONE forward whose consumers are all differentiated by one backward: each
    # consumer then replaces origins/directions by its pass-through outputs, so that the bundle's gradient travels back
    # along that chain through one buffer (fields._chained_positions).  Off for layouts a caller may evaluate and
    # differentiate several times independently.
    chain: bool = False

    @property
    def num_rays(self) -> int:
        return self.ebins.shape[0]

    @property
    def num_samples(self) -> int:
        return self.ebins.shape[1] - 1


class _Sliceable:
    """Minimal stand-in for the reference's TensorDataclass: tensor fields share leading batch dims."""

    def _map(self, fn):
        kw = {}
        for f in fields(self):
            v = getattr(self, f.name)
            if f.name == "_layout":
                kw[f.name] = None  # per-ray fast-path layout does not survive re-indexing
            elif torch.is_tensor(v):
                kw[f.name] = fn(v)
            elif isinstance(v, _Sliceable):
                kw[f.name] = v._map(fn)
            elif isinstance(v, dict):
                kw[f.name] = {k: (fn(t) if torch.is_tensor(t) else t) for k, t in v.items()}
        return replace(self, **kw)

    def to(self, device):
        return self._map(lambda t: t.to(device))


@dataclass
class Frustums(_Sliceable):
    """cameras/rays.py:32-103."""

    origins: Tensor  # [*bs,3]
    directions: Tensor  # [*bs,3]
    starts: Tensor  # [*bs,1]
    ends: Tensor  # [*bs,1]
    pixel_area: Tensor  # [*bs,1]
    offsets: Optional[Tensor] = None

    @property
    def shape(self):
        return torch.broadcast_shapes(self.origins.shape[:-1], self.starts.shape[:-1])

    def get_positions(self) -> Tensor:
        """cameras/rays.py:49-58 (plain torch: this is the generic, user-facing accessor)."""
        pos = self.origins + self.directions * (self.starts + self.ends) / 2
        if self.offsets is not None:
            pos = pos + self.offsets
        return pos

    def get_start_positions(self) -> Tensor:
        return self.origins + self.directions * self.starts


@dataclass
class RaySamples(_Sliceable):
    """cameras/rays.py:106-150."""

    frustums: Frustums
    camera_indices: Optional[Tensor] = None  # [*bs,1]
    deltas: Optional[Tensor] = None  # [*bs,1]
    spacing_starts: Optional[Tensor] = None
    spacing_ends: Optional[Tensor] = None
    spacing_to_euclidean_fn: Optional[Callable] = None
    metadata: Optional[Dict[str, Tensor]] = None
    times: Optional[Tensor] = None
    _layout: Optional[RayLayout] = field(default=None, repr=False, compare=False)

    def __getattribute__(self, name):
        # `deltas` of samples made by this package's samplers are formed on first use: the kernels take the bin
        # edges of the per-ray layout, so the [R,S,1] difference is only materialised for callers that read it
        v = object.__getattribute__(self, name)
        if v is None and name == "deltas":
            lay = object.__getattribute__(self, "_layout")
            if lay is not None:
                v = (lay.ebins[:, 1:] - lay.ebins[:, :-1])[..., None]
                object.__setattr__(self, "deltas", v)
        return v

    @property
    def shape(self):
        return self.frustums.shape

    def get_weights(self, densities: Tensor) -> Tensor:
        """alpha-composited sample weights, cameras/rays.py:128-150.  densities [R,S,1] -> [R,S,1]."""
        assert self.deltas is not None
        r, s = densities.shape[0], densities.shape[-2]
        deltas = self.deltas.expand(*densities.shape).reshape(-1, s)
        w = ops.sample_weights(densities.reshape(-1, s), deltas)
        return w.view(*densities.shape)


@dataclass
class RayBundle(_Sliceable):
    """cameras/rays.py:191-295."""

    origins: Tensor  # [*bs,3]
    directions: Tensor  # [*bs,3]
    pixel_area: Tensor  # [*bs,1]
    camera_indices: Optional[Tensor] = None  # [*bs,1]
    nears: Optional[Tensor] = None
    fars: Optional[Tensor] = None
    metadata: Dict[str, Tensor] = field(default_factory=dict)
    times: Optional[Tensor] = None

    def __len__(self) -> int:
        return torch.numel(self.origins) // self.origins.shape[-1]

    @property
    def shape(self):
        return self.origins.shape[:-1]

    def flatten(self) -> "RayBundle":
        return self._map(lambda t: t.reshape(-1, t.shape[-1]))

    def __getitem__(self, idx) -> "RayBundle":
        return self._map(lambda t: t[idx])

    def get_row_major_sliced_ray_bundle(self, start_idx: int, end_idx: int) -> "RayBundle":
        return self.flatten()[start_idx:end_idx]

    def set_camera_indices(self, camera_index: int) -> None:
        self.camera_indices = torch.ones_like(self.origins[..., 0:1]).long() * camera_index

    def get_ray_samples(self, bin_starts: Tensor, bin_ends: Tensor, spacing_starts: Optional[Tensor] = None,
                        spacing_ends: Optional[Tensor] = None,
                        spacing_to_euclidean_fn: Optional[Callable] = None) -> RaySamples:
        """cameras/rays.py:251-295: bin_starts/bin_ends [R,S,1] -> RaySamples with [R,S,*] views."""
        s = bin_starts.shape[-2]
        deltas = bin_ends - bin_starts
        expand = lambda t: None if t is None else t[..., None, :].expand(*t.shape[:-1], s, t.shape[-1])  # noqa: E731
        frustums = Frustums(origins=expand(self.origins), directions=expand(self.directions), starts=bin_starts,
                            ends=bin_ends, pixel_area=expand(self.pixel_area))
        return RaySamples(frustums=frustums, camera_indices=expand(self.camera_indices), deltas=deltas,
                          spacing_starts=spacing_starts, spacing_ends=spacing_ends,
                          spacing_to_euclidean_fn=spacing_to_euclidean_fn,
                          metadata={k: expand(v) for k, v in self.metadata.items()} if self.metadata else None,
                          times=expand(self.times))


def samples_from_layout(bundle: RayBundle, layout: RayLayout, spacing_to_euclidean_fn: Callable) -> RaySamples:
    """Public RaySamples whose tensors are views of the per-ray layout (no [R,S,3] materialisation)."""
    eb, sb = layout.ebins, layout.sbins
    s = eb.shape[1] - 1
    expand = lambda t: None if t is None else t[..., None, :].expand(*t.shape[:-1], s, t.shape[-1])  # noqa: E731
    frustums = Frustums(origins=expand(bundle.origins), directions=expand(bundle.directions), starts=eb[:, :-1, None],
                        ends=eb[:, 1:, None], pixel_area=expand(bundle.pixel_area))
    # as RayBundle.get_ray_samples (cameras/rays.py:251-295), with `deltas` left to RaySamples.__getattribute__
    return RaySamples(frustums=frustums, camera_indices=expand(bundle.camera_indices), deltas=None,
                      spacing_starts=sb[:, :-1, None], spacing_ends=sb[:, 1:, None],
                      spacing_to_euclidean_fn=spacing_to_euclidean_fn,
                      metadata={k: expand(v) for k, v in bundle.metadata.items()} if bundle.metadata else None,
                      times=expand(bundle.times), _layout=layout)
