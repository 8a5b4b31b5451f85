"""Ray samplers with the reference's class names and call signatures.

Mirrors `nerfstudio/model_components/ray_samplers.py` for the samplers thermal-nerfacto uses
(UniformLinDispPiecewiseSampler / UniformSampler, PDFSampler, ProposalNetworkSampler).

RNG: the reference draws its stratification jitter with `torch.rand` inside the samplers, one [R,1] tensor
per sampler call (single_jitter).  The same draws happen here, in the same order, on the ray device; tests
and parity runs may pass them explicitly (`jitter=` / `jitters=`) to reproduce a CPU reference run.
"""
from typing import Any, Callable, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import ops
from .rays import RayBundle, RayLayout, RaySamples, samples_from_layout


class Sampler(nn.Module):
    def __init__(self, num_samples: Optional[int] = None) -> None:
        super().__init__()
        self.num_samples = num_samples

    def generate_ray_samples(self, *args, **kwargs) -> Any:
        raise NotImplementedError

    def forward(self, *args, **kwargs) -> Any:
        return self.generate_ray_samples(*args, **kwargs)


class _PiecewiseSpacing:
    """spacing_to_euclidean_fn of the piecewise sampler (ray_samplers.py:115-116, 244-245) as a callable
    object, so that PDFSampler can recognise it and let the kernel do the conversion."""

    def __init__(self, nears: Tensor, fars: Tensor):
        self.nears, self.fars = nears, fars
        self._s = None  # spacing-domain near/far: only the torch path (__call__) needs them

    @property
    def s_near(self) -> Tensor:
        if self._s is None:
            self._s = (self.fn(self.nears), self.fn(self.fars))
        return self._s[0]

    @property
    def s_far(self) -> Tensor:
        if self._s is None:
            self._s = (self.fn(self.nears), self.fn(self.fars))
        return self._s[1]

    @staticmethod
    def fn(x: Tensor) -> Tensor:
        return torch.where(x < 1, x / 2, 1 - 1 / (2 * x))

    @staticmethod
    def fn_inv(x: Tensor) -> Tensor:
        return torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x))

    def __call__(self, x: Tensor) -> Tensor:
        return self.fn_inv(x * self.s_far + (1 - x) * self.s_near)


class SpacedSampler(Sampler):
    """ray_samplers.py:53-128 (generic spacing functions run as torch ops; the piecewise subclass is a kernel)."""

    def __init__(self, spacing_fn: Callable, spacing_fn_inv: Callable, num_samples: Optional[int] = None,
                 train_stratified=True, single_jitter=False) -> None:
        super().__init__(num_samples=num_samples)
        self.train_stratified = train_stratified
        self.single_jitter = single_jitter
        self.spacing_fn = spacing_fn
        self.spacing_fn_inv = spacing_fn_inv

    def _draw_jitter(self, num_rays: int, per_ray: int, device) -> Optional[Tensor]:
        if not (self.train_stratified and self.training):
            return None
        cols = 1 if self.single_jitter else per_ray
        return torch.rand((num_rays, cols), dtype=torch.float32, device=device)

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, num_samples: Optional[int] = None,
                             jitter: Optional[Tensor] = None) -> RaySamples:
        assert ray_bundle is not None and ray_bundle.nears is not None and ray_bundle.fars is not None
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        num_rays = ray_bundle.origins.shape[0]
        dev = ray_bundle.origins.device
        bins = torch.linspace(0.0, 1.0, num_samples + 1).to(dev)[None, ...]
        if jitter is None:
            jitter = self._draw_jitter(num_rays, num_samples + 1, dev)
        if jitter is not None:
            centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
            upper = torch.cat([centers, bins[..., -1:]], -1)
            lower = torch.cat([bins[..., :1], centers], -1)
            bins = lower + (upper - lower) * jitter
        s_near, s_far = (self.spacing_fn(x) for x in (ray_bundle.nears, ray_bundle.fars))

        def spacing_to_euclidean_fn(x):
            return self.spacing_fn_inv(x * s_far + (1 - x) * s_near)

        ebins = spacing_to_euclidean_fn(bins)
        bins = bins.expand(num_rays, -1)
        layout = RayLayout(ray_bundle.origins, ray_bundle.directions, ebins.contiguous(), bins.contiguous(),
                           ray_bundle.nears.reshape(-1), ray_bundle.fars.reshape(-1))
        return samples_from_layout(ray_bundle, layout, spacing_to_euclidean_fn)


class UniformSampler(SpacedSampler):
    """ray_samplers.py:131-152."""

    def __init__(self, num_samples: Optional[int] = None, train_stratified=True, single_jitter=False) -> None:
        super().__init__(num_samples=num_samples, spacing_fn=lambda x: x, spacing_fn_inv=lambda x: x,
                         train_stratified=train_stratified, single_jitter=single_jitter)


class UniformLinDispPiecewiseSampler(SpacedSampler):
    """ray_samplers.py:225-248: first half of the samples uniform up to distance 1, second half linear in
    disparity.  One kernel produces the spacing and euclidean bin edges."""

    def __init__(self, num_samples: Optional[int] = None, train_stratified=True, single_jitter=False) -> None:
        super().__init__(num_samples=num_samples, spacing_fn=_PiecewiseSpacing.fn,
                         spacing_fn_inv=_PiecewiseSpacing.fn_inv, train_stratified=train_stratified,
                         single_jitter=single_jitter)

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, num_samples: Optional[int] = None,
                             jitter: Optional[Tensor] = None) -> RaySamples:
        assert ray_bundle is not None and ray_bundle.nears is not None and ray_bundle.fars is not None
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        num_rays = ray_bundle.origins.shape[0]
        if jitter is None:
            jitter = self._draw_jitter(num_rays, num_samples + 1, ray_bundle.origins.device)
        sbins, ebins = ops.piecewise_bins(ray_bundle.nears, ray_bundle.fars, num_samples, jitter)
        layout = RayLayout(ray_bundle.origins, ray_bundle.directions, ebins, sbins, ray_bundle.nears.reshape(-1),
                           ray_bundle.fars.reshape(-1))
        return samples_from_layout(ray_bundle, layout, _PiecewiseSpacing(ray_bundle.nears, ray_bundle.fars))


class PDFSampler(Sampler):
    """ray_samplers.py:251-372: inverse-CDF resampling of the previous level's histogram."""

    def __init__(self, num_samples: Optional[int] = None, train_stratified: bool = True, single_jitter: bool = False,
                 include_original: bool = True, histogram_padding: float = 0.01) -> None:
        super().__init__(num_samples=num_samples)
        self.train_stratified = train_stratified
        self.include_original = include_original
        self.histogram_padding = histogram_padding
        self.single_jitter = single_jitter

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, ray_samples: Optional[RaySamples] = None,
                             weights: Optional[Tensor] = None, num_samples: Optional[int] = None, eps: float = 1e-5,
                             jitter: Optional[Tensor] = None, anneal: Optional[Tensor] = None) -> RaySamples:
        """`anneal` (extension): device scalar a; the histogram is then weights ** a, formed inside the kernel
        (ProposalNetworkSampler's annealing with an exponent a captured CUDA graph can follow)."""
        if ray_samples is None or ray_bundle is None:
            raise ValueError("ray_samples and ray_bundle must be provided")
        assert weights is not None, "weights must be provided"
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        assert ray_samples.spacing_starts is not None and ray_samples.spacing_ends is not None, \
            "ray_sample spacing_starts and spacing_ends must be provided"
        assert ray_samples.spacing_to_euclidean_fn is not None, "ray_samples.spacing_to_euclidean_fn must be provided"
        lay = getattr(ray_samples, "_layout", None)  # reference-built RaySamples carry no layout
        if lay is not None:
            existing_bins = lay.sbins
        else:
            existing_bins = torch.cat([ray_samples.spacing_starts[..., 0], ray_samples.spacing_ends[..., -1:, 0]],
                                      dim=-1).contiguous()
        num_rays = existing_bins.shape[0]
        if jitter is None and self.train_stratified and self.training:
            cols = 1 if self.single_jitter else num_samples + 1
            jitter = torch.rand((num_rays, cols), device=existing_bins.device)
        fn = ray_samples.spacing_to_euclidean_fn
        piecewise = isinstance(fn, _PiecewiseSpacing)
        nears = fn.nears if piecewise else torch.zeros(num_rays, device=existing_bins.device)
        fars = fn.fars if piecewise else torch.ones(num_rays, device=existing_bins.device)
        sbins, ebins = ops.pdf_sample(weights[..., 0], existing_bins, nears, fars, num_samples, jitter,
                                      self.histogram_padding, eps, anneal=anneal)
        if self.include_original:
            sbins, _ = torch.sort(torch.cat([existing_bins, sbins], -1), -1)
            ebins = fn(sbins)
        elif not piecewise:
            ebins = fn(sbins)
        layout = RayLayout(ray_bundle.origins, ray_bundle.directions, ebins, sbins,
                           nears.reshape(-1), fars.reshape(-1))
        return samples_from_layout(ray_bundle, layout, fn)


class ProposalNetworkSampler(Sampler):
    """ray_samplers.py:523-618."""

    def __init__(self, num_proposal_samples_per_ray: Tuple[int, ...] = (64,), num_nerf_samples_per_ray: int = 32,
                 num_proposal_network_iterations: int = 2, single_jitter: bool = False,
                 update_sched: Callable = lambda x: 1, initial_sampler: Optional[Sampler] = None,
                 pdf_sampler: Optional[PDFSampler] = None) -> None:
        super().__init__()
        self.num_proposal_samples_per_ray = num_proposal_samples_per_ray
        self.num_nerf_samples_per_ray = num_nerf_samples_per_ray
        self.num_proposal_network_iterations = num_proposal_network_iterations
        self.update_sched = update_sched
        if self.num_proposal_network_iterations < 1:
            raise ValueError("num_proposal_network_iterations must be >= 1")
        self.initial_sampler = initial_sampler if initial_sampler is not None \
            else UniformLinDispPiecewiseSampler(single_jitter=single_jitter)
        self.pdf_sampler = pdf_sampler if pdf_sampler is not None \
            else PDFSampler(include_original=False, single_jitter=single_jitter)
        self._anneal = 1.0
        self._steps_since_update = 0
        self._step = 0
        # device copy of the anneal exponent: the PDF kernel reads it at run time, so a CUDA graph captured at one
        # step follows the schedule on later replays (set_anneal rewrites it before the replay)
        self._anneal_dev: Optional[Tensor] = None
        # engine.GraphedTrainStep pins the "updated" decision while it captures a graph variant and applies the
        # bookkeeping itself after each replay
        self._forced_updated: Optional[bool] = None

    def set_anneal(self, anneal: float) -> None:
        self._anneal = anneal
        if self._anneal_dev is not None:
            self._anneal_dev.fill_(float(anneal))

    def _anneal_tensor(self, device) -> Tensor:
        if self._anneal_dev is None or self._anneal_dev.device != device:
            self._anneal_dev = torch.full((1,), float(self._anneal), dtype=torch.float32, device=device)
        return self._anneal_dev

    def step_cb(self, step):
        self._step = step
        self._steps_since_update += 1

    def will_update(self) -> bool:
        """The `updated` decision the next generate_ray_samples call takes (ray_samplers.py:591)."""
        return bool(self._steps_since_update > self.update_sched(self._step) or self._step < 10)

    def mark_sampled(self, updated: bool) -> None:
        """The bookkeeping generate_ray_samples does at its end (ray_samplers.py:612-613), for runners that replay
        a captured forward instead of calling it."""
        if updated:
            self._steps_since_update = 0

    @staticmethod
    def _density(density_fn: Callable, ray_samples: RaySamples) -> Tensor:
        """`density_fns[i](positions)` of the reference.  When the callable is a Field's bound `density_fn`
        the field is evaluated on the samples directly (positions are then formed inside the kernel from
        the per-ray layout instead of being materialised as an [R,S,3] tensor)."""
        owner = getattr(density_fn, "__self__", None)
        if owner is not None and getattr(density_fn, "__name__", "") == "density_fn" and hasattr(owner, "get_density"):
            return owner.get_density(ray_samples)[0]
        return density_fn(ray_samples.frustums.get_positions())

    def draw_jitters(self, num_rays: int, device) -> Optional[List[Optional[Tensor]]]:
        """The stratification draws generate_ray_samples would make, in its order (one per level), so that a caller
        can fix the random stream before reordering or overlapping the work that consumes it."""
        out: List[Optional[Tensor]] = []
        n = self.num_proposal_network_iterations
        for i_level in range(n + 1):
            s = self.num_proposal_samples_per_ray[i_level] if i_level < n else self.num_nerf_samples_per_ray
            smp = self.initial_sampler if i_level == 0 else self.pdf_sampler
            if not (smp.train_stratified and smp.training):
                out.append(None)
            else:
                out.append(torch.rand((num_rays, 1 if smp.single_jitter else s + 1), dtype=torch.float32, device=device))
        return None if all(j is None for j in out) else out

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None,
                             density_fns: Optional[List[Callable]] = None,
                             jitters: Optional[List[Tensor]] = None) -> Tuple[RaySamples, List, List]:
        assert ray_bundle is not None
        assert density_fns is not None
        weights_list = []
        ray_samples_list = []
        n = self.num_proposal_network_iterations
        weights = None
        ray_samples = None
        updated = self.will_update() if self._forced_updated is None else self._forced_updated
        on_device = ray_bundle.origins.is_cuda
        for i_level in range(n + 1):
            is_prop = i_level < n
            num_samples = self.num_proposal_samples_per_ray[i_level] if is_prop else self.num_nerf_samples_per_ray
            jit = None if jitters is None else jitters[i_level]
            if i_level == 0:
                ray_samples = self.initial_sampler(ray_bundle, num_samples=num_samples, jitter=jit)
            else:
                assert weights is not None
                if on_device and isinstance(self.pdf_sampler, PDFSampler):
                    # weights ** anneal (:602) inside the PDF kernel, exponent read from the device
                    ray_samples = self.pdf_sampler(ray_bundle, ray_samples, weights, num_samples=num_samples,
                                                   jitter=jit, anneal=self._anneal_tensor(ray_bundle.origins.device))
                else:
                    annealed_weights = weights if self._anneal == 1.0 else torch.pow(weights, self._anneal)
                    ray_samples = self.pdf_sampler(ray_bundle, ray_samples, annealed_weights,
                                                   num_samples=num_samples, jitter=jit)
            if is_prop:
                if updated:
                    density = self._density(density_fns[i_level], ray_samples)
                else:
                    with torch.no_grad():
                        density = self._density(density_fns[i_level], ray_samples)
                weights = ray_samples.get_weights(density)
                weights_list.append(weights)
                ray_samples_list.append(ray_samples)
        if self._forced_updated is None:
            self.mark_sampled(updated)
        assert ray_samples is not None
        return ray_samples, weights_list, ray_samples_list
