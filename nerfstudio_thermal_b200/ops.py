"""torch.autograd bindings of the libtn_b200 kernels (thin: shape checks, allocation, launch).

All tensors are CUDA float32; every launch goes to torch's current stream so the ops compose with
torch code, CUDA graphs and one-process-per-GPU data parallelism.  No op here has a CPU path.
"""
import os
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import call, float_array, ptr, ptr_array, stream

ACT_NONE, ACT_SIGMOID, ACT_TRUNC_EXP = 0, 1, 2
BG_NONE, BG_LAST_SAMPLE, BG_CONSTANT = 0, 1, 2


def _f32c(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------------- hash grid
class HashGridSpec:
    """Static description of one grid: per-level scales (host floats), L, F, log2 T."""

    def __init__(self, scalings: Sequence[float], features_per_level: int, log2_hashmap_size: int):
        self.scalings = [float(s) for s in scalings]
        self.num_levels = len(self.scalings)
        self.features = int(features_per_level)
        self.log2_T = int(log2_hashmap_size)
        self._c_scales = float_array(self.scalings)

    @property
    def out_dim(self) -> int:
        return self.num_levels * self.features

    @property
    def rows(self) -> int:
        return self.num_levels << self.log2_T


def hash_encode_indices(x: Tensor, spec: HashGridSpec) -> Tensor:
    """int32 [N, L, 8] table rows in the reference's corner order (test hook for bit-exactness)."""
    x = _f32c(x)
    n = x.shape[0]
    dummy = torch.zeros((spec.rows, spec.features), device=x.device)
    out = torch.empty((n, spec.out_dim), device=x.device)
    idx = torch.empty((n, spec.num_levels, 8), device=x.device, dtype=torch.int32)
    call("tn_hash_encode_fwd", ptr(x), ptr(dummy), 0, spec._c_scales, n, spec.num_levels, spec.features, spec.log2_T,
         0, ptr(out), ptr(idx), None, stream())
    return idx


# Encode forward keeps d(features)/dx of the FINE levels (scale >= 400: 6 of the 16 main-grid levels) in planar form
# for the backward, whose dL/dx of those levels becomes a streaming pre-pass instead of 8 latency-exposed gathers
# per level.  Measured (round 2, profiles/r02_jacobian.txt): backward 0.490 -> 0.445 ms/step, forward 0.231 -> 0.264,
# step 2.13 -> 2.07 ms.  (Round 1's all-levels, interleaved version lost: 75 MB of badly coalesced stores per call.)
SAVE_JACOBIAN = os.environ.get("TN_SAVE_JAC", "1") == "1"


class _HashEncodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, table, spec: HashGridSpec, table_f16, grad_sink, samples_per_ray):
        x = _f32c(x)
        n = x.shape[0]
        out = torch.empty((n, spec.out_dim), device=x.device, dtype=torch.float32)
        src, dtype = (table_f16, 1) if table_f16 is not None else (table, 0)
        # when dL/dx will be wanted, the forward also stores d(features)/dx (12*F bytes per point and level,
        # streamed) so that the backward does not gather the corner rows a second time
        jac = None
        if SAVE_JACOBIAN and ctx.needs_input_grad[0]:
            jac = torch.empty((spec.num_levels, spec.features * 3, n), device=x.device)  # only the fine levels are touched
        call("tn_hash_encode_fwd", ptr(x), ptr(src), dtype, spec._c_scales, n, spec.num_levels, spec.features,
             spec.log2_T, samples_per_ray, ptr(out), None, ptr(jac), stream(),
             tag=f"[L{spec.num_levels},T2^{spec.log2_T}{',jac' if jac is not None else ''}]", units=n)
        ctx.spec = spec
        ctx.samples_per_ray = samples_per_ray
        ctx.grad_sink = grad_sink
        ctx.save_for_backward(x, table, table_f16, jac)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, table, table_f16, jac = ctx.saved_tensors
        spec = ctx.spec
        dy = _f32c(dy)
        n = x.shape[0]
        sink = ctx.grad_sink
        want_table = ctx.needs_input_grad[1]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        if not want_table and dx is None:
            return None, None, None, None, None, None
        if want_table and sink is not None:
            dtable = sink  # scatter straight into the (flat) gradient buffer: no 64 MB temporary, no add pass
        else:  # fresh gradient (or scratch target when only dx is wanted: the kernel always scatters)
            dtable = torch.zeros_like(table, dtype=torch.float32)
        src, dtype = (table_f16, 1) if table_f16 is not None else (table, 0)
        call("tn_hash_encode_bwd", ptr(x), ptr(src), dtype, spec._c_scales, ptr(dy), n, spec.num_levels, spec.features,
             spec.log2_T, ctx.samples_per_ray, ptr(dtable), ptr(dx), ptr(jac if dx is not None else None), stream(),
             tag=f"[L{spec.num_levels},T2^{spec.log2_T}{',dx' if dx is not None else ''}]", units=n)
        return dx, (dtable if (want_table and sink is None) else None), None, None, None, None


def hash_encode(x: Tensor, table: Tensor, spec: HashGridSpec, table_f16: Optional[Tensor] = None,
                grad_sink: Optional[Tensor] = None, samples_per_ray: int = 0) -> Tensor:
    """x[N,3] in [0,1], table[L*T,F] -> [N, L*F].  field_components/encodings.py:420-461.

    grad_sink: optional float32 tensor shaped like `table` that the backward kernel accumulates the table
    gradient INTO (e.g. the parameter's slice of a FlatGradBuffer); autograd then receives no table gradient.
    samples_per_ray: S when x holds the [R,S] samples of R rays in ray-major order (a tiling hint, same results)."""
    return _HashEncodeFn.apply(x, table, spec, table_f16, grad_sink, int(samples_per_ray))


# ----------------------------------------------------------------------------------- positions
def _ray_grad_targets(g_o, g_d, like_o, like_d, zero: bool):
    """Where a ray-bundle consumer's backward puts d(origins), d(directions): onto the gradient handed down by the
    bundle's later consumers when there is one (returns accumulate=True), else into fresh buffers."""
    if g_o is not None and g_d is not None:
        return _f32c(g_o), _f32c(g_d), True
    make = torch.zeros_like if zero else torch.empty_like
    return make(like_o), make(like_d), False


def _ray_grad_finish(do, dd, g_o, g_d, accumulated: bool):
    if not accumulated:  # a lone upstream gradient (never on the model's path)
        do = do if g_o is None else do + g_o
        dd = dd if g_d is None else dd + g_d
    return do, dd


class _SamplePositionsFn(torch.autograd.Function):
    """chain=True also returns origins/directions as pass-through outputs: the bundle's NEXT consumer takes those,
    so that in the backward its gradient arrives here and this kernel adds onto it in place -- a bundle read by k
    consumers costs no zero-fills and no autograd additions (it used to cost 2k fills and 2(k-1) adds)."""

    @staticmethod
    def forward(ctx, origins, directions, ebins, chain):
        o, d, ebins = _f32c(origins), _f32c(directions), _f32c(ebins)
        r, s = ebins.shape[0], ebins.shape[1] - 1
        x = torch.empty((r * s, 3), device=ebins.device)
        sel = torch.empty((r * s,), device=ebins.device)
        call("tn_sample_positions_fwd", ptr(o), ptr(d), ptr(ebins), r, s, ptr(x), ptr(sel), stream())
        ctx.save_for_backward(o, d, ebins)
        ctx.mark_non_differentiable(sel)
        ctx.set_materialize_grads(False)
        return (x, sel, origins, directions) if chain else (x, sel)

    @staticmethod
    def backward(ctx, dx, _dsel, g_o=None, g_d=None):
        origins, directions, ebins = ctx.saved_tensors
        if dx is None:
            return g_o, g_d, None, None
        r, s = ebins.shape[0], ebins.shape[1] - 1
        do, dd, acc = _ray_grad_targets(g_o, g_d, origins, directions, zero=False)
        call("tn_sample_positions_bwd", ptr(origins), ptr(directions), ptr(ebins), ptr(_f32c(dx)), r, s, int(acc),
             ptr(do), ptr(dd), stream())
        do, dd = _ray_grad_finish(do, dd, g_o, g_d, acc)
        return do, dd, None, None


def sample_positions(origins: Tensor, directions: Tensor, ebins: Tensor, chain: bool = False):
    """Ray samples -> contracted, normalised grid coordinates x[R*S,3] and selector[R*S]
    (+ the pass-through origins/directions for the bundle's next consumer when chain=True)."""
    return _SamplePositionsFn.apply(origins, directions, ebins, chain)


class _ContractPointsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p):
        p = _f32c(p)
        n = p.shape[0]
        x = torch.empty_like(p)
        sel = torch.empty((n,), device=p.device)
        call("tn_contract_points_fwd", ptr(p), n, ptr(x), ptr(sel), stream())
        ctx.save_for_backward(p)
        ctx.mark_non_differentiable(sel)
        return x, sel

    @staticmethod
    def backward(ctx, dx, _dsel):
        (p,) = ctx.saved_tensors
        dp = torch.empty_like(p)
        call("tn_contract_points_bwd", ptr(p), ptr(_f32c(dx)), p.shape[0], ptr(dp), stream())
        return dp


def contract_points(p: Tensor) -> Tuple[Tensor, Tensor]:
    """World positions p[N,3] -> (x[N,3], selector[N])."""
    return _ContractPointsFn.apply(p)


# ----------------------------------------------------------------------------------- MLP
# "tc": tcgen05 tensor-core kernels (bf16x3 split precision, fp32 accumulate);  "simt": exact-fp32 FFMA kernels
MLP_BACKEND = {"fwd": "tc", "bwd": "tc"}


class _MlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, out_act, sinks, row_mul, out_scale, *wb):
        x = _f32c(x)
        row_mul = None if row_mul is None else _f32c(row_mul).view(-1)
        ctx.out_scale = float(out_scale)
        ws = [_f32c(t) for t in wb[0::2]]
        bs = [_f32c(t) for t in wb[1::2]]
        ctx.sinks = sinks
        n, x_stride = x.shape
        in_dim = ws[0].shape[1]  # x may carry padding columns beyond in_dim (a 63-wide input stored 64 wide)
        if x_stride < in_dim:
            raise ValueError(f"mlp: input has {x_stride} columns, first layer expects {in_dim}")
        width, out_dim, nl = ws[0].shape[0], ws[-1].shape[0], len(ws)
        y = torch.empty((n, out_dim), device=x.device)
        tag = f"[{in_dim}-{width}x{nl - 1}-{out_dim}]"
        mask = None
        if MLP_BACKEND["fwd"] == "tc":
            if MLP_BACKEND["bwd"] == "tc" and any(ctx.needs_input_grad):
                mask = torch.empty((n, nl - 1, max(1, width // 32)), device=x.device, dtype=torch.int32)
            call("tn_mlp_tc_fwd", ptr(x), n, in_dim, x_stride, width, out_dim, nl, ptr_array(ws), ptr_array(bs), out_act,
                 ptr(row_mul), ctx.out_scale, ptr(y), ptr(mask), stream(), tag=tag, units=n)
        else:
            assert row_mul is None and out_scale == 1.0, "output multipliers are a tensor-core epilogue (see mlp())"
            xd = x if x_stride == in_dim else x[:, :in_dim].contiguous()
            call("tn_mlp_fwd", ptr(xd), n, in_dim, width, out_dim, nl, ptr_array(ws), ptr_array(bs), out_act, ptr(y),
                 stream(), tag=tag, units=n)
        ctx.out_act = out_act
        ctx.nl = nl
        ctx.save_for_backward(x, mask, row_mul, *ws, *bs)
        return y

    @staticmethod
    def backward(ctx, dy):
        saved = ctx.saved_tensors
        x, mask, row_mul, nl = saved[0], saved[1], saved[2], ctx.nl
        ws, bs = list(saved[3:3 + nl]), list(saved[3 + nl:])
        n, x_stride = x.shape
        in_dim = ws[0].shape[1]
        width, out_dim = ws[0].shape[0], ws[-1].shape[0]
        sinks = ctx.sinks
        if sinks is not None:  # accumulate straight into the parameters' (flat-buffer) gradients
            dws, dbs = [s[0] for s in sinks], [s[1] for s in sinks]
        else:
            dws = [torch.zeros_like(w) for w in ws]
            dbs = [torch.zeros_like(b) for b in bs]
        dx = None
        if ctx.needs_input_grad[0]:
            # the tensor-core kernel writes whole 16/32/64-wide rows (zeros in the padding columns); any other
            # padding is never touched by a kernel and must already be zero
            tile_w = 16 if in_dim <= 16 else (32 if in_dim <= 32 else 64)
            covered = x_stride == in_dim or (MLP_BACKEND["bwd"] == "tc" and x_stride == tile_w)
            dx = torch.empty_like(x) if covered else torch.zeros_like(x)
        tag = f"[{in_dim}-{width}x{nl - 1}-{out_dim}]"
        if MLP_BACKEND["bwd"] == "tc":
            call("tn_mlp_tc_bwd", ptr(x), ptr(_f32c(dy)), ptr(mask), n, in_dim, x_stride, width, out_dim, nl,
                 ptr_array(ws), ptr_array(bs), ctx.out_act, ptr(row_mul), ctx.out_scale, ptr(dx), ptr_array(dws),
                 ptr_array(dbs), stream(), tag=tag, units=n)
        else:
            xd = x if x_stride == in_dim else x[:, :in_dim].contiguous()
            dxd = dx if x_stride == in_dim or dx is None else torch.empty_like(xd)
            call("tn_mlp_bwd", ptr(xd), ptr(_f32c(dy)), n, in_dim, width, out_dim, nl, ptr_array(ws), ptr_array(bs),
                 ctx.out_act, ptr(dxd), ptr_array(dws), ptr_array(dbs), stream(), tag=tag, units=n)
            if dx is not None and dxd is not dx:
                dx.zero_()
                dx[:, :in_dim] = dxd
        grads = []
        for dw, db in zip(dws, dbs):
            grads += [None, None] if sinks is not None else [dw, db]
        return (dx, None, None, None, None, *grads)


def mlp(x: Tensor, weights: List[Tensor], biases: List[Tensor], out_act: int = ACT_NONE, sinks=None,
        row_mul: Optional[Tensor] = None, out_scale: float = 1.0) -> Tensor:
    """Fully fused MLP (ReLU hidden activations).  field_components/mlp.py:159-178.

    sinks: optional [(dW_i, db_i)] float32 tensors the backward kernel accumulates INTO (the parameters' slices of
    a FlatGradBuffer); autograd then receives no weight gradients.
    row_mul [N] / out_scale: the activated output rows are multiplied by out_scale * row_mul (no gradient to
    row_mul) -- the density epilogue `average_init_density * trunc_exp(.) * selector` of the fields."""
    wb = []
    for w, b in zip(weights, biases):
        wb += [w, b]
    if MLP_BACKEND["fwd"] == "tc" and MLP_BACKEND["bwd"] == "tc":
        return _MlpFn.apply(x, out_act, sinks, row_mul, out_scale, *wb)
    y = _MlpFn.apply(x, out_act, sinks, None, 1.0, *wb)
    if row_mul is not None:
        y = y * row_mul.detach().view(-1, 1)
    return y if out_scale == 1.0 else y * out_scale


def mlp_shape_supported(in_dim: int, width: int, out_dim: int, n_layers: int) -> bool:
    return n_layers in (2, 3) and 1 <= in_dim <= 64 and width in (16, 64) and 1 <= out_dim <= 16


def sh4(directions: Tensor) -> Tensor:
    """16 SH components of d[N,3] (no gradient, as in the reference: encodings.py:792 is @torch.no_grad)."""
    d = _f32c(directions.detach())
    out = torch.empty((d.shape[0], 16), device=d.device)
    call("tn_sh4", ptr(d), d.shape[0], ptr(out), stream())
    return out


# ----------------------------------------------------------------------------------- samplers
_LINSPACE_CACHE = {}


def _host_linspace(key, make, device):
    """Constants the reference computes with torch.linspace on the fly; computed once on the HOST
    (bit-identical to the reference's values) and cached per device."""
    k = (key, str(device))
    if k not in _LINSPACE_CACHE:
        _LINSPACE_CACHE[k] = make().to(device)
    return _LINSPACE_CACHE[k]


def _jitter_arg(jitter: Optional[Tensor], rays: int, per_ray: int):
    """jitter None | [R,1] (single jitter) | [R,per_ray] -> (contiguous tensor | None, per-sample flag)."""
    if jitter is None:
        return None, 0
    j = _f32c(jitter)
    if j.numel() == rays:
        return j.view(-1), 0
    if j.numel() == rays * per_ray:
        return j.view(-1), 1
    raise ValueError(f"jitter has {j.numel()} elements; expected {rays} or {rays * per_ray}")


def piecewise_bins(nears: Tensor, fars: Tensor, num_samples: int, jitter: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """UniformLinDispPiecewiseSampler bins: (spacing bins, euclidean bins), both [R, S+1]."""
    nears, fars = _f32c(nears).view(-1), _f32c(fars).view(-1)
    r = nears.shape[0]
    unit = _host_linspace(("unit", num_samples), lambda: torch.linspace(0.0, 1.0, num_samples + 1), nears.device)
    sb = torch.empty((r, num_samples + 1), device=nears.device)
    eb = torch.empty_like(sb)
    jit, per_sample = _jitter_arg(jitter, r, num_samples + 1)
    call("tn_piecewise_bins", ptr(unit), ptr(nears), ptr(fars), ptr(jit), per_sample, r, num_samples, ptr(sb), ptr(eb),
         stream())
    return sb, eb


def pdf_sample(weights: Tensor, sbins_old: Tensor, nears: Tensor, fars: Tensor, num_samples: int,
               jitter: Optional[Tensor], histogram_padding: float = 0.01, eps: float = 1e-5,
               anneal: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """PDFSampler (include_original=False): new (spacing bins, euclidean bins) [R, S_new+1].
    anneal: optional device scalar; the histogram is then weights ** anneal (evaluated in the kernel)."""
    w = _f32c(weights.detach())
    r, s_old = w.shape
    nb = num_samples + 1
    if jitter is None:
        u = _host_linspace(("u_eval", nb), lambda: torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb) + 1.0 / (2 * nb),
                           w.device)
    else:
        u = _host_linspace(("u_train", nb), lambda: torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb), w.device)
    sb = torch.empty((r, nb), device=w.device)
    eb = torch.empty_like(sb)
    jit, per_sample = _jitter_arg(jitter, r, nb)
    call("tn_pdf_sample", ptr(w), ptr(_f32c(sbins_old)), ptr(_f32c(nears).view(-1)), ptr(_f32c(fars).view(-1)), ptr(u),
         ptr(jit), per_sample, ptr(anneal), r, s_old, num_samples, histogram_padding, eps, ptr(sb), ptr(eb), stream())
    return sb, eb


# ----------------------------------------------------------------------------------- rendering
class _WeightsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sigma, deltas):
        sigma, deltas = _f32c(sigma), _f32c(deltas)
        r, s = sigma.shape
        w = torch.empty_like(sigma)
        call("tn_weights_fwd", ptr(sigma), ptr(deltas), r, s, ptr(w), stream())
        ctx.save_for_backward(sigma, deltas)
        return w

    @staticmethod
    def backward(ctx, dw):
        sigma, deltas = ctx.saved_tensors
        r, s = sigma.shape
        ds = torch.empty_like(sigma)
        call("tn_weights_bwd", ptr(sigma), ptr(deltas), ptr(_f32c(dw)), r, s, ptr(ds), stream())
        return ds, None


def sample_weights(sigma: Tensor, deltas: Tensor) -> Tensor:
    """sigma, deltas [R,S] -> weights [R,S].  cameras/rays.py:128-150.  (deltas carry no gradient.)"""
    return _WeightsFn.apply(sigma, deltas.detach())


class _RenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, colour, starts, ends, bg_mode, bg, eval_mode, want_depth, bins):
        weights = _f32c(weights)
        r, s = weights.shape
        c = 0 if colour is None else colour.shape[-1]
        dev = weights.device
        colour_c = None if colour is None else _f32c(colour).view(r, s, c)
        if bins is not None:  # starts = bins[:, :-1], ends = bins[:, 1:] read in place (row stride S+1)
            bins = _f32c(bins)
            assert bins.shape == (r, s + 1)
            starts_c = ends_c = None
            p_st, p_en, ts = ptr(bins), ptr(bins) + 4, s + 1
        else:
            starts_c = None if starts is None else _f32c(starts).view(r, s)
            ends_c = None if ends is None else _f32c(ends).view(r, s)
            p_st, p_en, ts = ptr(starts_c), ptr(ends_c), 0
        rgb = torch.empty((r, c), device=dev) if c else None
        acc = torch.empty((r, 1), device=dev)
        med = exp = minmax = None
        if want_depth:
            med = torch.empty((r, 1), device=dev)
            exp = torch.empty((r, 1), device=dev)
            # cached device constant + clone: no host->device copy at call time (CUDA-graph capture safe)
            minmax = _host_linspace(("minmax",), lambda: torch.tensor([float("inf"), float("-inf")]), dev).clone()
        bg_arr = float_array(bg) if bg is not None else None
        call("tn_render_fwd", ptr(weights), ptr(colour_c), p_st, p_en, r, s, ts, c, bg_mode, bg_arr,
             int(eval_mode), ptr(rgb), ptr(acc), ptr(med), ptr(exp), ptr(minmax), stream())
        ctx.bg_mode, ctx.bg, ctx.c = bg_mode, bg, c
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(weights, colour_c, starts_c, ends_c, bins)
        outs = (rgb, acc, med, exp, minmax)
        ctx.mark_non_differentiable(*[o for o in (med, minmax) if o is not None])
        return outs

    @staticmethod
    def backward(ctx, d_rgb, d_acc, _d_med, d_exp, _d_mm):
        weights, colour, starts, ends, bins = ctx.saved_tensors
        r, s = weights.shape
        c = ctx.c
        if bins is not None:
            p_st, p_en, ts = ptr(bins), ptr(bins) + 4, s + 1
        else:
            p_st, p_en, ts = ptr(starts), ptr(ends), 0
        dw = torch.empty_like(weights) if ctx.needs_input_grad[0] else None
        dcol = torch.empty_like(colour) if (c and ctx.needs_input_grad[1]) else None
        bg_arr = float_array(ctx.bg) if ctx.bg is not None else None
        d_rgb = None if (d_rgb is None or not c) else _f32c(d_rgb)
        d_acc = None if d_acc is None else _f32c(d_acc)
        d_exp = None if d_exp is None else _f32c(d_exp)
        call("tn_render_bwd", ptr(weights), ptr(colour), p_st, p_en, ptr(d_rgb), ptr(d_acc), ptr(d_exp), r,
             s, ts, c, ctx.bg_mode, bg_arr, ptr(dw), ptr(dcol), stream())
        return dw, dcol, None, None, None, None, None, None, None


def render(weights: Tensor, colour: Optional[Tensor], starts: Optional[Tensor], ends: Optional[Tensor], *,
           bg_mode: int = BG_NONE, bg: Optional[Sequence[float]] = None, eval_mode: bool = False,
           want_depth: bool = False, bins: Optional[Tensor] = None):
    """Fused renderer reductions over one launch.  `bins` [R,S+1] may replace starts/ends (= bins[:, :-1] and
    bins[:, 1:], read in place).

    Returns (rgb [R,C] | None, accumulation [R,1], median depth [R,1] | None,
             UNCLIPPED expected depth [R,1] | None, (min,max) of sample midpoints [2] | None).
    """
    return _RenderFn.apply(weights, colour, starts, ends, bg_mode, None if bg is None else tuple(bg), eval_mode,
                           want_depth, bins)


def library_info() -> str:
    lib = _lib.load()
    return f"libtn_b200 v{lib.tn_version()} ({lib.tn_build_arch().decode()}) at {_lib.LIB_PATH}"
