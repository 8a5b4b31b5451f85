"""torch.autograd bindings of the glue fusions and per-ray loss kernels (csrc/tn_fused.cu)."""
from typing import Optional, Tuple

import ctypes

import torch
from torch import Tensor

from ._lib import call, float_array, ptr, ptr_array, stream
from .ops import _f32c
from .ops import _ray_grad_finish as ops_ray_grad_finish
from .ops import _ray_grad_targets as ops_ray_grad_targets


class _FieldSplitFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, sel, sh, emb_ray, rays, samples, geo_dim, scale):
        h, sel, sh = _f32c(h), _f32c(sel), _f32c(sh)
        emb_ray = None if emb_ray is None else _f32c(emb_ray)
        emb_dim = 0 if emb_ray is None else emb_ray.shape[-1]
        n = rays * samples
        density = torch.empty((n,), device=h.device)
        in_dim = 16 + geo_dim + emb_dim
        x_stride = (in_dim + 3) // 4 * 4  # rows padded to whole 16-byte runs (63 -> 64) for the MLP's tile loads
        x = torch.empty((n, x_stride), device=h.device)
        call("tn_field_split_fwd", ptr(h), ptr(sel), ptr(sh), ptr(emb_ray), rays, samples, h.shape[-1], geo_dim, emb_dim,
             x_stride, float(scale), ptr(density), ptr(x), stream())
        ctx.dims = (rays, samples, geo_dim, emb_dim, float(scale), x_stride)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(h, sel)
        return density, x

    @staticmethod
    def backward(ctx, d_density, dx):
        h, sel = ctx.saved_tensors
        rays, samples, geo_dim, emb_dim, scale, x_stride = ctx.dims
        dh = torch.empty_like(h)
        want_emb = emb_dim > 0 and ctx.needs_input_grad[3] and dx is not None
        demb = torch.empty((rays, emb_dim), device=h.device) if want_emb else None
        call("tn_field_split_bwd", ptr(h), ptr(sel), ptr(None if d_density is None else _f32c(d_density)),
             ptr(None if dx is None else _f32c(dx)), rays, samples, h.shape[-1], geo_dim, emb_dim, x_stride, scale,
             ptr(dh), ptr(demb), stream())
        return dh, None, None, demb, None, None, None, None


def field_split(h: Tensor, sel: Tensor, sh: Tensor, emb_ray: Optional[Tensor], rays: int, samples: int, geo_dim: int,
                scale: float) -> Tuple[Tensor, Tensor]:
    """h[R*S,1+geo] -> (density[R*S], head input [R*S, pad4(16+geo+emb)], zero padded: `ops.mlp` takes the padded
    rows as they are).  fields/nerfacto_field.py:221-228, 335-344."""
    return _FieldSplitFn.apply(h, sel, sh, emb_ray, rays, samples, geo_dim, scale)


class LaunchScratch:
    """Small device buffers a module keeps ACROSS steps for kernels that leave them in their initial state themselves
    (the ticket / accumulator words of tn_ray_heads_fwd and tn_density_l1, the per-ray buffer tn_field_head_bwd
    accumulates into and tn_embed_bwd clears): no fill, clone or zeros launch around those calls.  A buffer is only
    ever created outside CUDA-graph capture (a tensor first made during capture holds nothing until a replay); get()
    returns None then and the caller takes the self-initialising path for that call."""

    def __init__(self) -> None:
        self._bufs = {}
        self.dirty = set()  # keys whose buffer a failed / partial backward may have left non-zero

    def get(self, key, make):
        t = self._bufs.get(key)
        if t is None:
            if (torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()) \
                    or torch.is_inference_mode_enabled():
                return None
            # never evicted: a captured CUDA graph holds the buffer's ADDRESS, not a reference (callers only ask for
            # small buffers: a few words per call site, <= 8 MB per distinct batch size)
            t = self._bufs[key] = make()
        return t


def heads_scratch(scratch: Optional[LaunchScratch], device, slot) -> Optional[Tensor]:
    """{+inf, -inf, 0, 0, ticket 0}: the launch-wide words of one tn_ray_heads_fwd call site (`slot`: call sites that
    may run concurrently -- the two branches -- use different words)."""
    if scratch is None:
        return None
    return scratch.get(("heads", str(device), slot),
                       lambda: torch.tensor([float("inf"), float("-inf"), 0.0, 0.0, 0.0]).to(device))


class _FieldHeadFn(torch.autograd.Function):
    """A NerfactoField's colour head with its input assembly as one autograd node: ONE tensor-core kernel each way
    (tn_field_head_fwd / _bwd).  Neither the 63-wide head input nor its gradient ever reaches HBM.

    emb_weight / cam_idx given: `emb_ray` are rows of that table (ray_features) and the backward turns the head's
    per-ray first-layer gradient straight into the table's gradient (tn_embed_bwd: no [R,32] gradient, no matmul, no
    index_add_), into `emb_sink` when there is one."""

    @staticmethod
    def forward(ctx, h, sel, sh, emb_ray, emb_weight, cam_idx, emb_sink, scratch, rays, samples, scale, out_act, sinks,
                *wb):
        h, sel, sh, emb_ray = _f32c(h), _f32c(sel), _f32c(sh), _f32c(emb_ray)
        ws = [_f32c(t) for t in wb[0::2]]
        bs = [_f32c(t) for t in wb[1::2]]
        n, out_dim = rays * samples, ws[-1].shape[0]
        density = torch.empty((n,), device=h.device)
        y = torch.empty((n, out_dim), device=h.device)
        need_mask = any(ctx.needs_input_grad)
        mask = torch.empty((n, 2, 2), device=h.device, dtype=torch.int32) if need_mask else None
        call("tn_field_head_fwd", ptr(h), ptr(sel), ptr(sh), ptr(emb_ray), rays, samples, out_dim, float(scale),
             ptr_array(ws), ptr_array(bs), out_act, ptr(density), ptr(y), ptr(mask), stream(), tag=f"[63-64x2-{out_dim}]",
             units=n)
        ctx.dims = (rays, samples, float(scale), out_dim, out_act)
        ctx.sinks, ctx.emb_sink, ctx.scratch = sinks, emb_sink, scratch
        ctx.table = emb_weight is not None and cam_idx is not None
        ctx.table_rows = emb_weight.shape[0] if ctx.table else 0
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(h, sel, sh, emb_ray, mask, cam_idx if ctx.table else None, *ws, *bs)
        return density, y

    @staticmethod
    def backward(ctx, d_density, dy):
        saved = ctx.saved_tensors
        h, sel, sh, emb_ray, mask, cam_idx = saved[:6]
        ws, bs = list(saved[6:9]), list(saved[9:12])
        rays, samples, scale, out_dim, out_act = ctx.dims
        if dy is None:
            dy = torch.zeros((rays * samples, out_dim), device=h.device)
        sinks = ctx.sinks
        if sinks is not None:
            dws, dbs = [s_[0] for s_ in sinks], [s_[1] for s_ in sinks]
        else:
            dws, dbs = [torch.zeros_like(w) for w in ws], [torch.zeros_like(b) for b in bs]
        dh = torch.empty_like(h)
        want_table = ctx.table and ctx.needs_input_grad[4]
        # dz1_ray: accumulated into by the head kernel; with the table path a persistent buffer that tn_embed_bwd clears
        key = ("dz1", str(h.device), rays)
        keep = ctx.scratch.get(key, lambda: torch.zeros((rays, 64), device=h.device)) \
            if (want_table and ctx.scratch is not None and rays <= 32768) else None
        if keep is not None:
            if key in ctx.scratch.dirty:
                keep.zero_()
            ctx.scratch.dirty.add(key)
            dz1_ray = keep
        else:
            dz1_ray = torch.zeros((rays, 64), device=h.device)
        call("tn_field_head_bwd", ptr(_f32c(dy)), ptr(mask), ptr(h), ptr(sel), ptr(sh), ptr(emb_ray),
             ptr(None if d_density is None else _f32c(d_density)), rays, samples, out_dim, scale, ptr_array(ws),
             ptr_array(bs), out_act, ptr(dh), ptr(dz1_ray), ptr_array(dws), ptr_array(dbs), stream(),
             tag=f"[63-64x2-{out_dim}]", units=rays * samples)
        demb = dtable = None
        if want_table:
            esink = ctx.emb_sink
            dtab = esink if esink is not None else torch.zeros((ctx.table_rows, 32), device=h.device)
            call("tn_embed_bwd", ptr(dz1_ray), ptr(ws[0]), ptr(cam_idx), rays, 64, ws[0].shape[1], 31, 32,
                 int(keep is not None), ptr(dtab), stream())
            if keep is not None:
                ctx.scratch.dirty.discard(key)
            dtable = None if esink is not None else dtab
        elif ctx.needs_input_grad[3]:
            demb = dz1_ray @ ws[0][:, 31:63]
        grads = []
        for dw, db in zip(dws, dbs):
            grads += [None, None] if sinks is not None else [dw, db]
        return (dh, None, None, demb, dtable, None, None, None, None, None, None, None, None, *grads)


def field_head(h: Tensor, sel: Tensor, sh: Tensor, emb_ray: Tensor, rays: int, samples: int, geo_dim: int,
               scale: float, weights, biases, out_act: int, sinks=None, emb_weight: Optional[Tensor] = None,
               cam_idx: Optional[Tensor] = None, emb_sink: Optional[Tensor] = None,
               scratch: Optional[LaunchScratch] = None) -> Tuple[Tensor, Tensor]:
    """(density[R*S], head output [R*S, C]) of a NerfactoField from the density-MLP output h[R*S,16]: trunc_exp *
    selector, the [SH | geo | appearance] concatenation and the 63-64-64-C colour head.
    fields/nerfacto_field.py:221-228, 335-348.
    emb_weight [num_cameras,32] + cam_idx [R] (int64): emb_ray = emb_weight[cam_idx] (ray_features); the embedding's
    gradient is then formed by the backward itself (see _FieldHeadFn)."""
    wb = []
    for w, b in zip(weights, biases):
        wb += [w, b]
    assert geo_dim == 15 and emb_ray.shape[-1] == 32 and h.shape[-1] == 16, "see field_head_supported"
    if emb_weight is not None:
        assert emb_weight.shape[-1] == 32 and cam_idx is not None and cam_idx.dtype == torch.int64
        cam_idx = cam_idx.contiguous()
    return _FieldHeadFn.apply(h, sel, sh, emb_ray, emb_weight, cam_idx, emb_sink, scratch, rays, samples, scale,
                              out_act, sinks, *wb)


@torch.no_grad()
def ray_features(directions: Tensor, emb_weight: Optional[Tensor] = None,
                 cam_idx: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor]]:
    """(SH basis [R,16] of the normalised directions (d+1)/2, appearance-embedding rows [R,E] | None) of a ray batch in
    one launch.  fields/base_field.py:136-142 + encodings.py:792-795 (under no_grad in the reference too) +
    field_components/embedding.py:48-55 (values only: the gradient is _FieldHeadFn's)."""
    d = _f32c(directions.detach())
    r = d.shape[0]
    sh = torch.empty((r, 16), device=d.device)
    emb = w = None
    e = 0
    if emb_weight is not None:
        w = _f32c(emb_weight.detach())
        e = w.shape[-1]
        emb = torch.empty((r, e), device=d.device)
        cam_idx = cam_idx.contiguous()
    call("tn_ray_features", ptr(d), ptr(w), ptr(cam_idx if emb is not None else None), r, e, ptr(sh), ptr(emb),
         stream())
    return sh, emb


def field_head_supported(h_width: int, geo_dim: int, emb_dim: int, samples: int, head) -> bool:
    """Shapes the fused head backward handles: 16-wide density-MLP output, a 64-wide 3-layer head whose input
    (16 + geo + emb) fills the 64-column tile, and at most 8 rays per 128-point tile."""
    from . import ops
    return (h_width == 16 and geo_dim == 15 and emb_dim == 32 and samples >= 19
            and len(head.layers) == 3 and head.layers[0].weight.shape[0] == 64 and head.layers[1].weight.shape == (64, 64)
            and head.layers[2].weight.shape[0] <= 16 and head._out_act is not None
            and ops.MLP_BACKEND == {"fwd": "tc", "bwd": "tc"})


class _DensityActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, sel, scale):
        h, sel = _f32c(h), _f32c(sel).view(-1)
        n, width = h.shape
        out = torch.empty((n,), device=h.device)
        call("tn_density_act_fwd", ptr(h), width, ptr(sel), n, float(scale), ptr(out), stream())
        ctx.scale = float(scale)
        ctx.save_for_backward(h, sel)
        return out

    @staticmethod
    def backward(ctx, dd):
        h, sel = ctx.saved_tensors
        n, width = h.shape
        dh = torch.zeros_like(h) if width > 1 else torch.empty_like(h)
        call("tn_density_act_bwd", ptr(h), width, ptr(sel), ptr(_f32c(dd).view(-1)), n, ctx.scale, ptr(dh), width,
             stream())
        return dh, None, None


def density_act(h: Tensor, sel: Tensor, scale: float) -> Tensor:
    """scale * trunc_exp(h[:, 0]) * selector -> [N] for an MLP output h[N, width] whose column 0 is the raw
    density.  fields/density_fields.py:116-117, fields/nerfacto_field.py:227-228."""
    return _DensityActFn.apply(h, sel, scale)


class _PropDensityFn(torch.autograd.Function):
    """ray samples -> proposal density, one kernel each way (csrc/tn_prop.cu).  chain: see ops._SamplePositionsFn."""

    @staticmethod
    def forward(ctx, origins, directions, ebins, table, w1, b1, w2, b2, spec, scale, grad_sink, mlp_sinks, chain):
        ctx.mlp_sinks = mlp_sinks
        o, d, ebins = _f32c(origins), _f32c(directions), _f32c(ebins)
        w1, b1, w2, b2 = _f32c(w1), _f32c(b1), _f32c(w2), _f32c(b2)
        r, s = ebins.shape[0], ebins.shape[1] - 1
        density = torch.empty((r * s,), device=ebins.device)
        call("tn_prop_density_fwd", ptr(o), ptr(d), ptr(ebins), ptr(table), spec._c_scales, r, s,
             spec.num_levels, spec.log2_T, w1.shape[0], ptr(w1), ptr(b1), ptr(w2), ptr(b2), float(scale), ptr(density),
             stream(), tag=f"[L{spec.num_levels},S{s}]", units=r * s)
        ctx.spec, ctx.scale, ctx.grad_sink = spec, float(scale), grad_sink
        ctx.save_for_backward(o, d, ebins, table, w1, b1, w2, b2)
        ctx.set_materialize_grads(False)
        return (density, origins, directions) if chain else density

    @staticmethod
    def backward(ctx, d_density, g_o=None, g_d=None):
        origins, directions, ebins, table, w1, b1, w2, b2 = ctx.saved_tensors
        if d_density is None:
            return (g_o, g_d) + (None,) * 11
        spec = ctx.spec
        r, s = ebins.shape[0], ebins.shape[1] - 1
        need_rays = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        d_o = d_d = None
        acc = False
        if need_rays:  # the kernel adds with atomics: onto the later consumers' gradient when there is one
            d_o, d_d, acc = ops_ray_grad_targets(g_o, g_d, origins, directions, zero=True)
        sink = ctx.grad_sink if ctx.needs_input_grad[3] else None
        dtable = sink if sink is not None else torch.zeros_like(table)
        ms = ctx.mlp_sinks
        if ms is not None:
            (dw1, db1), (dw2, db2) = ms
        else:
            dw1, db1, dw2, db2 = (torch.zeros_like(t) for t in (w1, b1, w2, b2))
        call("tn_prop_density_bwd", ptr(origins), ptr(directions), ptr(ebins), ptr(table), spec._c_scales, r, s,
             spec.num_levels, spec.log2_T, w1.shape[0], ptr(w1), ptr(b1), ptr(w2), ptr(b2), ctx.scale,
             ptr(_f32c(d_density)), ptr(dtable), ptr(dw1), ptr(db1), ptr(dw2), ptr(db2), ptr(d_o), ptr(d_d), stream(),
             tag=f"[L{spec.num_levels},S{s}{',dx' if need_rays else ''}]", units=r * s)
        if need_rays:
            d_o, d_d = ops_ray_grad_finish(d_o, d_d, g_o, g_d, acc)
        else:
            d_o, d_d = g_o, g_d
        dt = dtable if (ctx.needs_input_grad[3] and sink is None) else None
        if ms is not None:
            dw1 = db1 = dw2 = db2 = None
        return d_o, d_d, None, dt, dw1, db1, dw2, db2, None, None, None, None, None


def prop_density(origins: Tensor, directions: Tensor, ebins: Tensor, table: Tensor, w1: Tensor, b1: Tensor, w2: Tensor,
                 b2: Tensor, spec, scale: float, grad_sink: Optional[Tensor] = None, mlp_sinks=None,
                 chain: bool = False):
    """HashMLPDensityField.get_density for ray samples, fused end to end.  fields/density_fields.py:95-118.
    grad_sink / mlp_sinks: optional accumulation targets for the table and [(dW1,db1),(dW2,db2)] gradients.
    chain=True: returns (density, origins, directions), the last two pass-through for the bundle's next consumer."""
    return _PropDensityFn.apply(origins, directions, ebins, table, w1, b1, w2, b2, spec, scale, grad_sink, mlp_sinks,
                                chain)


class _EmbedRowsFn(torch.autograd.Function):
    """weight[idx] whose backward is an index_add_ (atomics) instead of torch's sort-based embedding backward."""

    @staticmethod
    def forward(ctx, weight, idx):
        ctx.save_for_backward(idx)
        ctx.shape = weight.shape
        return weight.index_select(0, idx)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        gw = torch.zeros(ctx.shape, device=g.device, dtype=g.dtype)
        gw.index_add_(0, idx, g.contiguous())
        return gw, None


def embed_rows(weight: Tensor, idx: Tensor) -> Tensor:
    """Embedding lookup for a few thousand per-ray indices (field_components/embedding.py:48-55)."""
    return _EmbedRowsFn.apply(weight, idx)


class _CameraOptFn(torch.autograd.Function):
    """CameraOptimizer.apply_to_raybundle (SO3xR3 / shared) as one kernel each way (csrc/tn_model.cu).
    sink: optional accumulation target for d(pose) (the parameter's slice of a FlatGradBuffer; the kernel adds with
    atomics): no zeros launch, and autograd has nothing to accumulate."""

    @staticmethod
    def forward(ctx, pose, frozen, cam, origins, directions, shared, sink):
        pose, origins, directions = _f32c(pose), _f32c(origins), _f32c(directions)
        cam = cam.contiguous()
        r = origins.shape[0]
        oo, dd = torch.empty_like(origins), torch.empty_like(directions)
        call("tn_camera_opt_fwd", ptr(pose), ptr(frozen), ptr(cam), ptr(origins), ptr(directions), r, int(shared),
             ptr(oo), ptr(dd), stream())
        ctx.shared, ctx.sink = int(shared), sink
        ctx.save_for_backward(pose, frozen, cam, directions)
        ctx.set_materialize_grads(False)
        return oo, dd

    @staticmethod
    def backward(ctx, g_o, g_d):
        pose, frozen, cam, directions = ctx.saved_tensors
        if g_o is None and g_d is None:
            return None, None, None, None, None, None, None
        sink = ctx.sink
        dpose = sink if sink is not None else torch.zeros_like(pose)
        call("tn_camera_opt_bwd", ptr(pose), ptr(frozen), ptr(cam), ptr(directions),
             ptr(None if g_o is None else _f32c(g_o)), ptr(None if g_d is None else _f32c(g_d)), directions.shape[0],
             ctx.shared, ptr(dpose), stream())
        # rays are data: the reference's bundle tensors do not require grad either
        return (None if sink is not None else dpose), None, None, None, None, None, None


def camera_opt_apply(pose: Tensor, frozen: Optional[Tensor], camera_indices: Tensor, origins: Tensor,
                     directions: Tensor, shared: bool, sink: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """(origins + t_c, R(w_c) directions) for the per-camera pose adjustments.
    cameras/camera_optimizers.py:132-176, cameras/lie_groups.py:24-59."""
    return _CameraOptFn.apply(pose, frozen, camera_indices, origins, directions, shared, sink)


class _PixelLossesFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, thermal, image, is_thermal):
        rgb, image, is_thermal = _f32c(rgb), _f32c(image), _f32c(is_thermal)
        thermal = None if thermal is None else _f32c(thermal).view(-1)
        losses = torch.empty((4,), device=rgb.device)
        call("tn_pixel_losses", ptr(rgb), ptr(thermal), ptr(image), ptr(is_thermal), rgb.shape[0], None, ptr(losses),
             None, None, stream())
        ctx.has_thermal = thermal is not None
        ctx.save_for_backward(rgb, thermal, image, is_thermal)
        ctx.set_materialize_grads(False)
        return tuple(losses.unbind(0))  # four scalars: indexing a vector would cost a zeros+copy per term backward

    @staticmethod
    def backward(ctx, *gs):
        rgb, thermal, image, is_thermal = ctx.saved_tensors
        d_rgb = torch.empty_like(rgb)
        d_th = torch.empty_like(thermal) if ctx.has_thermal else None
        zero = _zero_scalar(rgb.device)
        g = torch.stack([zero if t is None else t.reshape(()) for t in gs])
        call("tn_pixel_losses", ptr(rgb), ptr(thermal), ptr(image), ptr(is_thermal), rgb.shape[0], ptr(_f32c(g)), None,
             ptr(d_rgb), ptr(d_th), stream())
        return d_rgb, (None if d_th is None else d_th.view(-1, 1)), None, None


_ZERO = {}


def _zero_scalar(device) -> Tensor:
    k = str(device)
    if k not in _ZERO:
        _ZERO[k] = torch.zeros((), device=device)
    return _ZERO[k]


def pixel_losses(rgb: Tensor, thermal: Optional[Tensor], image: Tensor, is_thermal: Tensor) -> Tuple[Tensor, ...]:
    """(rgb MSE, thermal MSE, tv_pixel, cross_channel) scalars (un-multiplied) for a patch-ordered batch.
    models/thermal_nerfacto.py:286-354; rgb[R,3], thermal[R,1] or None, image[R,3], is_thermal[R]."""
    return _PixelLossesFn.apply(rgb, thermal, image, is_thermal)


_L1_PARTIALS = 64


class _DensityL1Fn(torch.autograd.Function):
    """Value in the forward (per-CTA partial sums; the CTA that finishes last adds them up when the caller lends a
    ticket word), the four gradients in the backward, already multiplied by the upstream scalar: one launch each way."""

    @staticmethod
    def forward(ctx, d, d2, dt, d2t, mult, rgb_mult, scratch):
        ctx.shape = d.shape
        d, d2, dt, d2t = (_f32c(t).view(-1) for t in (d, d2, dt, d2t))
        n = d.numel()
        if rgb_mult == 1:  # symmetric branch (:331-334)
            vm, m, rm = mult, mult, mult
        else:              # asymmetric stop-gradient pattern (:336-344)
            vm, m, rm = mult * (1.0 + rgb_mult), mult, mult * rgb_mult
        partial = torch.empty((_L1_PARTIALS + 1,), device=d.device)
        ticket = None if scratch is None else scratch.get(
            ("l1", str(d.device)), lambda: torch.zeros((1,), device=d.device, dtype=torch.int32))
        call("tn_density_l1", ptr(d), ptr(d2), ptr(dt), ptr(d2t), n, float(vm), float(m), float(rm), None, ptr(partial),
             _L1_PARTIALS, ptr(ticket), None, None, None, None, stream())
        ctx.mults = (float(vm), float(m), float(rm))
        ctx.save_for_backward(d, d2, dt, d2t)
        return partial[_L1_PARTIALS] if ticket is not None else partial[:_L1_PARTIALS].sum()

    @staticmethod
    def backward(ctx, g):
        d, d2, dt, d2t = ctx.saved_tensors
        s = ctx.shape
        grads = [torch.empty_like(d) for _ in range(4)]
        call("tn_density_l1", ptr(d), ptr(d2), ptr(dt), ptr(d2t), d.numel(), *ctx.mults, ptr(_f32c(g)), None,
             _L1_PARTIALS, None, ptr(grads[0]), ptr(grads[1]), ptr(grads[2]), ptr(grads[3]), stream())
        return (*[t.view(s) for t in grads], None, None, None)


def density_l1(d: Tensor, d2: Tensor, dt: Tensor, d2t: Tensor, mult: float, rgb_mult: float,
               scratch: Optional[LaunchScratch] = None) -> Tensor:
    """density_loss of ThermalNerfactoModel.get_loss_dict (models/thermal_nerfacto.py:328-344): the value in one
    launch, the gradients to all four densities in one launch."""
    return _DensityL1Fn.apply(d, d2, dt, d2t, mult, rgb_mult, scratch)


class _DistortionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w, sbins):
        w, sbins = _f32c(w), _f32c(sbins)
        r, s = w.shape
        loss_ray = torch.empty((r,), device=w.device)
        dw = torch.empty_like(w) if ctx.needs_input_grad[0] else None
        call("tn_distortion_loss", ptr(w), ptr(sbins), r, s, ptr(loss_ray), ptr(dw), stream())
        ctx.save_for_backward(dw)
        ctx.rays = r
        return loss_ray.sum() / r

    @staticmethod
    def backward(ctx, g):
        (dw,) = ctx.saved_tensors
        return (None if dw is None else dw * (g / ctx.rays)), None


def distortion_loss_rays(w: Tensor, sbins: Tensor) -> Tensor:
    """mean over rays of lossfun_distortion(sbins, w).  model_components/losses.py:139-158."""
    return _DistortionFn.apply(w, sbins.detach())


class _InterlevelFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w_fine, c_fine, w_prop, c_prop):
        w_fine, c_fine, w_prop, c_prop = _f32c(w_fine), _f32c(c_fine), _f32c(w_prop), _f32c(c_prop)
        r, sf = w_fine.shape
        sp = w_prop.shape[1]
        loss_ray = torch.empty((r,), device=w_fine.device)
        dwp = torch.empty_like(w_prop) if ctx.needs_input_grad[2] else None
        call("tn_interlevel_loss", ptr(w_fine), ptr(c_fine), ptr(w_prop), ptr(c_prop), r, sf, sp, ptr(loss_ray), ptr(dwp),
             stream())
        ctx.save_for_backward(dwp)
        ctx.count = r * sf
        return loss_ray.sum() / ctx.count

    @staticmethod
    def backward(ctx, g):
        (dwp,) = ctx.saved_tensors
        return None, None, (None if dwp is None else dwp * (g / ctx.count)), None


def interlevel_loss_level(w_fine: Tensor, c_fine: Tensor, w_prop: Tensor, c_prop: Tensor) -> Tensor:
    """mean(lossfun_outer(c_fine, w_fine, c_prop, w_prop)); the fine histogram carries no gradient.
    model_components/losses.py:87-103, 117-135."""
    return _InterlevelFn.apply(w_fine.detach(), c_fine.detach(), w_prop, c_prop.detach())


class _CameraRegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pose, trans_pen, rot_pen, scale, sink):
        pose = _f32c(pose)
        out = torch.empty((3,), device=pose.device)
        call("tn_camera_reg_fwd", ptr(pose), pose.shape[0], float(trans_pen), float(rot_pen), float(scale), ptr(out),
             stream())
        ctx.coef = (float(trans_pen), float(rot_pen), float(scale))
        ctx.sink = sink
        ctx.save_for_backward(pose)
        ctx.set_materialize_grads(False)
        reg, tn_, rn_ = out.unbind(0)
        ctx.mark_non_differentiable(tn_, rn_)
        return reg, tn_, rn_

    @staticmethod
    def backward(ctx, g, _gt, _gr):
        (pose,) = ctx.saved_tensors
        if g is None:
            return None, None, None, None, None
        sink = ctx.sink
        dpose = sink if sink is not None else torch.empty_like(pose)
        call("tn_camera_reg_bwd", ptr(pose), ptr(_f32c(g)), pose.shape[0], *ctx.coef, int(sink is not None), ptr(dpose),
             stream())
        return (None if sink is not None else dpose), None, None, None, None


def camera_regularizer(pose: Tensor, trans_pen: float, rot_pen: float, scale: float,
                       sink: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """(regulariser, |translations|_F, |rotations|_F) of a CameraOptimizer's pose table in one launch.
    cameras/camera_optimizers.py:188-194, 200-204.  sink: accumulation target of d(pose) (see camera_opt_apply)."""
    return _CameraRegFn.apply(pose, trans_pen, rot_pen, scale, sink)


_SCALE_CACHE = {}


class _LossSumFn(torch.autograd.Function):
    """Dictionary values (sums of scale_k * term_k over the terms of an entry) and their total in one launch; the
    backward hands every term its slice of ONE vector product instead of a chain of scalar multiplies."""

    @staticmethod
    def forward(ctx, scales, slots, *terms):
        terms = [_f32c(t) for t in terms]
        k, n_slots = len(terms), max(slots) + 1
        dev = terms[0].device
        out = torch.empty((1 + n_slots,), device=dev)
        call("tn_loss_sum", ptr_array(terms), float_array(scales), (ctypes.c_int * k)(*slots), k, n_slots, ptr(out),
             stream())
        key = (tuple(scales), str(dev))
        if key not in _SCALE_CACHE:
            _SCALE_CACHE[key] = torch.tensor(scales, dtype=torch.float32).to(dev)
        ctx.scales, ctx.slots = _SCALE_CACHE[key], slots
        ctx.set_materialize_grads(False)
        return tuple(out.unbind(0))

    @staticmethod
    def backward(ctx, g_total, *g_slot):
        gv = (ctx.scales * g_total).unbind(0) if g_total is not None else [None] * len(ctx.slots)
        outs = []
        for i, j in enumerate(ctx.slots):  # an entry differentiated on its own adds its share
            t, g = gv[i], g_slot[j]
            if g is not None:
                t = ctx.scales[i] * g if t is None else t + ctx.scales[i] * g
            outs.append(t)
        return (None, None, *outs)


def loss_sum(entries) -> Tuple[Tensor, dict]:
    """entries: [(name, scalar term, scale)], names may repeat -> (total, {name: sum of scale * term}).
    models/thermal_nerfacto.py:284-388 multiplies every loss term by its config weight and engine/trainer.py:479
    adds the dictionary up -- two tiny launches per term each way; here one launch each way."""
    names = []
    for n, _, _ in entries:
        if n not in names:
            names.append(n)
    out = _LossSumFn.apply(tuple(float(s) for _, _, s in entries), tuple(names.index(n) for n, _, _ in entries),
                           *[t for _, t, _ in entries])
    return out[0], dict(zip(names, out[1:]))


# ----------------------------------------------------------------------------------- one launch per sampling level
_HEADS_INIT = {}


def _heads_scratch(device) -> Tensor:
    """[+inf, -inf, 0, 0]: the (min, max) of the sample midpoints and the two loss accumulators of a ray_heads
    launch, cloned from a cached device constant (no host->device copy: CUDA-graph capture safe)."""
    k = str(device)
    if k not in _HEADS_INIT:
        _HEADS_INIT[k] = torch.tensor([float("inf"), float("-inf"), 0.0, 0.0]).to(device)
    return _HEADS_INIT[k].clone()


@torch.no_grad()
def level_resample(sigma: Tensor, ebins: Tensor, sbins: Tensor, nears: Tensor, fars: Tensor, num_samples: int,
                   jitter: Optional[Tensor], anneal: Optional[Tensor] = None, histogram_padding: float = 0.01,
                   eps: float = 1e-5, want_depth: bool = True):
    """One proposal level after its density: get_weights -> median depth -> annealed PDF resampling in ONE launch
    (cameras/rays.py:128-150, renderers.py:547-557, ray_samplers.py:301-372 / :602).  The sampling chain carries no
    gradient in the reference (ray_samplers.py:360 detaches the bins), so this runs outside autograd.
    sigma [R,S] -> (weights [R,S], median depth [R,1] | None, sbins_new, ebins_new [R,num_samples+1])."""
    from . import ops

    sigma, ebins, sbins = _f32c(sigma), _f32c(ebins), _f32c(sbins)
    r, s = sigma.shape
    nb = num_samples + 1
    dev = sigma.device
    if jitter is None:
        u = ops._host_linspace(("u_eval", nb), lambda: torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb) + 1.0 / (2 * nb),
                               dev)
    else:
        u = ops._host_linspace(("u_train", nb), lambda: torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb), dev)
    jit, per_sample = ops._jitter_arg(jitter, r, nb)
    w = torch.empty_like(sigma)
    med = torch.empty((r, 1), device=dev) if want_depth else None
    sb = torch.empty((r, nb), device=dev)
    eb = torch.empty_like(sb)
    call("tn_level_resample", ptr(sigma), ptr(ebins), ptr(sbins), ptr(_f32c(nears).view(-1)), ptr(_f32c(fars).view(-1)),
         ptr(u), ptr(jit), per_sample, ptr(anneal), r, s, num_samples, float(histogram_padding), float(eps), ptr(w),
         ptr(med), ptr(sb), ptr(eb), stream(), tag=f"[S{s}->{num_samples}]", units=r * s)
    return w, med, sb, eb


class _RayHeadsFn(torch.autograd.Function):
    """Final level of a branch as ONE autograd node: get_weights + all renderers + distortion + interlevel loss
    forward (tn_ray_heads_fwd), and the ray-level backward of the whole branch -- final level and every proposal
    level -- in one launch (tn_ray_heads_bwd)."""

    _FIXED = 12  # positional arguments before *prop_sigma

    @staticmethod
    def forward(ctx, sigma, colour, ebins, sbins, bg_mode, bg, eval_mode, want_losses, prop_ebins, prop_sbins,
                prop_weights, scratch, *prop_sigma):
        sigma, colour, ebins, sbins = _f32c(sigma), _f32c(colour), _f32c(ebins), _f32c(sbins)
        r, s = sigma.shape
        c = colour.shape[-1]
        colour = colour.view(r, s, c)
        dev = sigma.device
        n_prop = len(prop_weights) if want_losses else 0
        prop_weights = [_f32c(t).view(r, -1) for t in prop_weights[:n_prop]]
        prop_sbins = [_f32c(t) for t in prop_sbins[:n_prop]]
        prop_S = [t.shape[1] for t in prop_weights]
        need_prop_grad = [want_losses and i < len(prop_sigma) and ctx.needs_input_grad[_RayHeadsFn._FIXED + i]
                          for i in range(n_prop)]
        prop_dw = [torch.empty_like(prop_weights[i]) if need_prop_grad[i] else None for i in range(n_prop)]
        w = torch.empty_like(sigma)
        rgb = torch.empty((r, c), device=dev)
        acc = torch.empty((r, 1), device=dev)
        med = torch.empty((r, 1), device=dev)
        exp = torch.empty((r, 1), device=dev)
        dw_dist = torch.empty_like(sigma) if (want_losses and ctx.needs_input_grad[0]) else None
        bg_arr = float_array(bg) if bg is not None else None
        ints = (ctypes.c_int * max(n_prop, 1))(*(prop_S or [0]))
        if scratch is not None:
            # the launch's last CTA writes the FINAL extrema, loss means and clipped depth and re-arms `scratch`
            out4 = torch.empty((4,), device=dev)
            minmax, loss_acc = out4[:2], out4[2:]
            exp_clip = torch.empty((r, 1), device=dev)
        else:
            out4 = _heads_scratch(dev)  # accumulated into by the kernel; scaled / used for the clip below
            minmax, loss_acc = out4[:2], out4[2:]
            exp_clip = None
        call("tn_ray_heads_fwd", ptr(sigma), ptr(colour), ptr(ebins), ptr(sbins), r, s, c, bg_mode, bg_arr,
             int(eval_mode), n_prop, ptr_array(prop_weights) if n_prop else None,
             ptr_array(prop_sbins) if n_prop else None, ints, ptr_array(prop_dw) if n_prop else None, ptr(w), ptr(rgb),
             ptr(acc), ptr(med), ptr(exp), ptr(minmax), ptr(loss_acc) if want_losses else None, ptr(dw_dist),
             ptr(scratch), 1.0 / r, 1.0 / (r * s), ptr(exp_clip), stream(),
             tag=f"[S{s},C{c},+{n_prop}]" if n_prop else f"[S{s},C{c}]", units=r * s)
        ctx.cfg = (bg_mode, bg, c, [i for i in range(n_prop) if need_prop_grad[i]], prop_S)
        ctx.set_materialize_grads(False)
        live = [i for i in range(n_prop) if need_prop_grad[i]]
        ctx.n_live = len(live)
        ctx.save_for_backward(sigma, colour, ebins, w, dw_dist, exp, minmax,
                              *[_f32c(prop_sigma[i]).view(r, -1) for i in live],
                              *[_f32c(prop_ebins[i]) for i in live], *[prop_dw[i] for i in live])
        if scratch is None:
            exp_clip = torch.clamp(exp, minmax[0], minmax[1])  # batch-global clip, renderers.py:574
        dist = inter = None
        if want_losses:  # means over rays / over rays x fine samples (losses.py:135, :158)
            dist, inter = (loss_acc if scratch is not None else loss_acc * _loss_scales(dev, r, s)).unbind(0)
        ctx.mark_non_differentiable(w, med, minmax)
        return rgb, acc, med, exp_clip, minmax, w, dist, inter

    @staticmethod
    def backward(ctx, d_rgb, d_acc, _d_med, d_exp, _d_mm, _d_w, g_dist, g_inter):
        bg_mode, bg, c, live, prop_S = ctx.cfg
        saved = ctx.saved_tensors
        sigma, colour, ebins, w, dw_dist, exp_raw, minmax = saved[:7]
        k = ctx.n_live
        p_sigma, p_ebins, p_dw = saved[7:7 + k], saved[7 + k:7 + 2 * k], saved[7 + 2 * k:7 + 3 * k]
        r, s = sigma.shape
        if d_exp is not None:  # the clip passes the gradient where it did not bite (torch.clamp's rule)
            d_exp = d_exp * ((exp_raw >= minmax[0]) & (exp_raw <= minmax[1])).to(d_exp.dtype)
        dsigma = torch.empty_like(sigma)
        dcol = torch.empty_like(colour) if ctx.needs_input_grad[1] else None
        p_dsigma = [torch.empty_like(t) for t in p_sigma]
        bg_arr = float_array(bg) if bg is not None else None
        f = lambda t: None if t is None else _f32c(t)  # noqa: E731
        ints = (ctypes.c_int * max(k, 1))(*([prop_S[i] for i in live] or [0]))
        call("tn_ray_heads_bwd", ptr(sigma), ptr(colour), ptr(ebins), ptr(w), ptr(dw_dist), ptr(f(d_rgb)), ptr(f(d_acc)),
             ptr(f(d_exp)), ptr(f(g_dist)), ptr(f(g_inter)), r, s, c, bg_mode, bg_arr, k,
             ptr_array(list(p_sigma)) if k else None, ptr_array(list(p_ebins)) if k else None,
             ptr_array(list(p_dw)) if k else None, ints, ptr(dsigma), ptr(dcol), ptr_array(p_dsigma) if k else None,
             stream(), tag=f"[S{s},C{c},+{k}]", units=r * s)
        grads = [None] * len(ctx.needs_input_grad[_RayHeadsFn._FIXED:])
        for j, i in enumerate(live):
            grads[i] = p_dsigma[j]
        return (dsigma, dcol, *([None] * (_RayHeadsFn._FIXED - 2)), *grads)


_LOSS_SCALES = {}


def _loss_scales(device, rays: int, samples: int) -> Tensor:
    k = (str(device), rays, samples)
    if k not in _LOSS_SCALES:
        _LOSS_SCALES[k] = torch.tensor([1.0 / rays, 1.0 / (rays * samples)], dtype=torch.float32).to(device)
    return _LOSS_SCALES[k]


def ray_heads(sigma: Tensor, colour: Tensor, ebins: Tensor, sbins: Tensor, *, bg_mode: int, bg, eval_mode: bool,
              want_losses: bool, prop_sigma=(), prop_ebins=(), prop_sbins=(), prop_weights=(),
              scratch: Optional[Tensor] = None):
    """sigma [R,S], colour [R,S,C], bin edges [R,S+1] of the final level (+ per proposal level: density [R,S_i] --
    the differentiable input --, bin edges and the weights tn_level_resample made of it)  ->
    (rgb [R,C], accumulation [R,1], median depth [R,1], expected depth [R,1] clipped to the launch-wide extrema of the
    sample midpoints (renderers.py:574), those (min,max) [2], weights [R,S] (no gradient: the losses that read them
    are inside), mean distortion loss | None, mean interlevel loss summed over the proposal levels | None).
    scratch: the call site's `heads_scratch` words; with them the kernel's last CTA finishes the extrema, the loss
    means and the clip itself (otherwise: a clone before and two small torch launches after the kernel)."""
    return _RayHeadsFn.apply(sigma, colour, ebins, sbins, bg_mode, None if bg is None else tuple(bg), eval_mode,
                             want_losses, tuple(prop_ebins), tuple(prop_sbins), tuple(prop_weights), scratch,
                             *prop_sigma)
