"""GPU parity tests for the glue fusions and the per-ray loss kernels (csrc/tn_fused.cu) against the oracle."""
import pytest
import torch

import oracle
from oracle import model as om
from oracle import sampling as osamp

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import nerfstudio_thermal_b200 as tn
    from nerfstudio_thermal_b200 import fused_ops, losses

DEV = "cuda"


def close(a, b, atol=1e-6, rtol=1e-5):
    torch.testing.assert_close(a.detach().cpu(), b.detach().cpu(), atol=atol, rtol=rtol)


@pytest.mark.parametrize("emb_dim", [32, 0])
def test_field_split_fwd_bwd(emb_dim):
    torch.manual_seed(0)
    R, S, geo = 13, 48, 15
    h = (torch.randn(R * S, 16) * 2).requires_grad_(True)
    sel = (torch.rand(R * S) > 0.2).float()
    sh = torch.randn(R, 16)
    emb = torch.randn(R, emb_dim).requires_grad_(True) if emb_dim else None
    # torch composition of fields/nerfacto_field.py:221-228 and :335-344
    dens_ref = 0.7 * oracle.trunc_exp(h[:, :1]) * sel[:, None]
    parts = [sh[:, None, :].expand(R, S, 16).reshape(-1, 16), h[:, 1:]]
    if emb_dim:
        parts.append(emb[:, None, :].expand(R, S, emb_dim).reshape(-1, emb_dim))
    x_ref = torch.cat(parts, -1)
    gd, gx = torch.randn(R * S), torch.randn_like(x_ref)
    ((dens_ref[:, 0] * gd).sum() + (x_ref * gx).sum()).backward()

    hg = h.detach().to(DEV).requires_grad_(True)
    eg = emb.detach().to(DEV).requires_grad_(True) if emb_dim else None
    dens, x = fused_ops.field_split(hg, sel.to(DEV), sh.to(DEV), eg, R, S, geo, 0.7)
    close(dens, dens_ref[:, 0], 1e-6, 1e-5)
    in_dim = x_ref.shape[1]
    assert x.shape[1] == (in_dim + 3) // 4 * 4  # rows padded to whole 16-byte runs for the MLP's tile loads
    assert torch.equal(x[:, :in_dim].cpu(), x_ref.detach())
    assert torch.count_nonzero(x[:, in_dim:]) == 0
    ((dens * gd.to(DEV)).sum() + (x[:, :in_dim] * gx.to(DEV)).sum()).backward()
    close(hg.grad, h.grad, 1e-5, 1e-5)
    if emb_dim:
        close(eg.grad, emb.grad, 1e-5, 1e-5)


@pytest.mark.parametrize("width", [1, 16])
def test_density_act(width):
    torch.manual_seed(1)
    n = 1000
    h = (torch.randn(n, width) * 6).requires_grad_(True)  # includes |x| > 15 after scaling: clamp in the gradient
    with torch.no_grad():
        h[:5, 0] = torch.tensor([20.0, -20.0, 15.0, -15.0, 0.0])
    sel = (torch.rand(n) > 0.3).float()
    ref = 1.3 * oracle.trunc_exp(h[:, 0]) * sel
    g = torch.randn(n)
    (ref * g).sum().backward()
    hg = h.detach().to(DEV).requires_grad_(True)
    out = fused_ops.density_act(hg, sel.to(DEV), 1.3)
    close(out, ref, 1e-6, 2e-6)
    (out * g.to(DEV)).sum().backward()
    close(hg.grad, h.grad, 1e-6, 2e-6)


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_losses_golden(golden, mode):
    """distortion + interlevel on the reference's sampling chain: values vs the reference, gradients vs autograd
    of the oracle's torch expressions"""
    g = golden("sampling.npz")

    def sdist(tag):
        return torch.cat([g[f"{mode}_{tag}_spacing_starts"][..., 0], g[f"{mode}_{tag}_spacing_ends"][..., -1:, 0]], -1)

    c0, c1, c2 = sdist("s0"), sdist("s1"), sdist("s2")
    w0 = g[f"{mode}_w0"][..., 0].clone().requires_grad_(True)
    w1 = g[f"{mode}_w1"][..., 0].clone().requires_grad_(True)
    w2 = g[f"{mode}_w2"][..., 0].clone().requires_grad_(True)
    dist = fused_ops.distortion_loss_rays(w2.detach().to(DEV).requires_grad_(True), c2.to(DEV))
    close(dist, g[f"{mode}_distortion"], 1e-7, 1e-5)
    w2g = w2.detach().to(DEV).requires_grad_(True)
    fused_ops.distortion_loss_rays(w2g, c2.to(DEV)).backward()
    close(w2g.grad, g[f"{mode}_ddist_w2"][..., 0], 1e-7, 1e-4)
    w0g, w1g = w0.detach().to(DEV).requires_grad_(True), w1.detach().to(DEV).requires_grad_(True)
    inter = (fused_ops.interlevel_loss_level(w2.detach().to(DEV), c2.to(DEV), w0g, c0.to(DEV))
             + fused_ops.interlevel_loss_level(w2.detach().to(DEV), c2.to(DEV), w1g, c1.to(DEV)))
    close(inter, g[f"{mode}_interlevel"], 1e-7, 1e-5)
    inter.backward()
    close(w0g.grad, g[f"{mode}_dinter_w0"][..., 0], 1e-7, 1e-4)
    close(w1g.grad, g[f"{mode}_dinter_w1"][..., 0], 1e-7, 1e-4)


def test_losses_random_vs_oracle():
    torch.manual_seed(3)
    R = 50
    for sf, sp in ((48, 96), (48, 256), (7, 33)):
        c = torch.sort(torch.rand(R, sf + 1), -1).values
        cp = torch.sort(torch.rand(R, sp + 1), -1).values
        w = torch.rand(R, sf) * 0.05
        wp = (torch.rand(R, sp) * 0.02).requires_grad_(True)
        sf_ = osamp.OracleSamples(None, None, None, None, None, c[:, :-1, None], c[:, 1:, None], None, None)
        sp_ = osamp.OracleSamples(None, None, None, None, None, cp[:, :-1, None], cp[:, 1:, None], None, None)
        ref = om.interlevel_loss([wp[..., None], w[..., None]], [sp_, sf_])
        ref.backward()
        wpg = wp.detach().to(DEV).requires_grad_(True)
        got = fused_ops.interlevel_loss_level(w.to(DEV), c.to(DEV), wpg, cp.to(DEV))
        close(got, ref, 1e-8, 1e-5)
        got.backward()
        close(wpg.grad, wp.grad, 1e-8, 1e-4)
        wv = w.clone().requires_grad_(True)
        refd = om.distortion_loss([wv[..., None]], [sf_])
        refd.backward()
        wg = w.to(DEV).requires_grad_(True)
        gotd = fused_ops.distortion_loss_rays(wg, c.to(DEV))
        close(gotd, refd, 1e-8, 1e-5)
        gotd.backward()
        close(wg.grad, wv.grad, 1e-8, 1e-4)


@pytest.mark.parametrize("S,max_res", [(256, 128), (96, 256), (48, 64)])
def test_fused_proposal_field_vs_oracle(S, max_res):
    """positions -> contraction -> 5-level grid -> 10-16-1 MLP -> trunc_exp*selector in one kernel, fwd + bwd"""
    torch.manual_seed(4)
    R = 70
    field = tn.HashMLPDensityField(torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), hidden_dim=16, num_levels=5,
                                   max_res=max_res, log2_hashmap_size=12,
                                   spatial_distortion=tn.SceneContraction(order=float("inf")))
    with torch.no_grad():
        field.encoding.hash_table.mul_(500.0)
    sd = {f"p.{k}": v.detach().clone().requires_grad_(v.dtype == torch.float32 and k != "aabb")
          for k, v in field.state_dict().items()}
    sd["p.mlp_base.0.hash_table"] = sd["p.encoding.hash_table"]
    field = field.to(DEV)
    o = (torch.randn(R, 3) * 0.4).requires_grad_(True)
    d = torch.nn.functional.normalize(torch.randn(R, 3), dim=-1).requires_grad_(True)
    nears, fars = torch.full((R, 1), 0.05), torch.full((R, 1), 1000.0)
    og, dg = o.detach().to(DEV).requires_grad_(True), d.detach().to(DEV).requires_grad_(True)
    rb = tn.RayBundle(origins=og, directions=dg, pixel_area=torch.ones(R, 1, device=DEV), nears=nears.to(DEV),
                      fars=fars.to(DEV))
    jit = torch.rand(R, 1)
    rs = tn.UniformLinDispPiecewiseSampler(single_jitter=True).train()(rb, num_samples=S, jitter=jit.to(DEV))
    assert rs._layout is not None
    dens = field.get_density(rs)[0]
    # oracle on the same samples
    samples = osamp.initial_samples(o, d, None, nears, fars, S, jit)
    ref = oracle.proposal_density(sd, "p", osamp.sample_positions(samples), num_levels=5, base_res=16, max_res=max_res,
                                  log2_hashmap_size=12)
    close(dens, ref, 1e-5, 2e-5)
    g = torch.rand_like(ref)
    # the interlevel hinge sends exact zeros for whole stretches of a ray: the backward kernel skips warps whose 32
    # samples all carry a zero upstream gradient -- rays 10..39 entirely, and ragged tails of the others
    g[10:40] = 0.0
    g[40:, S // 2 + 3:] = 0.0
    (ref * g).sum().backward()
    (dens * g.to(DEV)).sum().backward()

    def rel(a, b):
        return ((a.detach().cpu().double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()

    assert rel(field.encoding.hash_table.grad, sd["p.encoding.hash_table"].grad) < 1e-4
    for i in (0, 1):
        assert rel(field.mlp_base[1].layers[i].weight.grad, sd[f"p.mlp_base.1.layers.{i}.weight"].grad) < 1e-4
        assert rel(field.mlp_base[1].layers[i].bias.grad, sd[f"p.mlp_base.1.layers.{i}.bias"].grad) < 1e-4
    assert rel(og.grad, o.grad) < 1e-3 and rel(dg.grad, d.grad) < 1e-3
    # and the unfused kernel chain gives the same density
    field.fuse = False
    close(field.get_density(rs)[0], dens, 1e-6, 1e-5)


@pytest.mark.parametrize("shared", [False, True])
def test_camera_optimizer_kernel(shared):
    torch.manual_seed(5)
    R, cams = 257, 8
    pose = (torch.randn(1 if shared else cams, 6) * 0.05)
    pose[0 if shared else 3, 3:] = 0.0  # |w| below the 1e-4 clamp: constant-angle branch
    pose = pose.requires_grad_(True)
    frozen = torch.tensor([False] * 5 + [True] * 3)
    idx = torch.randint(0, cams, (R, 1))
    o, d = torch.randn(R, 3), torch.nn.functional.normalize(torch.randn(R, 3), dim=-1)
    src = pose.expand(cams, 6) if shared else pose
    ro, rd = om.apply_camera_optimizer(src, frozen, idx, o, d)
    go, gd = torch.randn(R, 3), torch.randn(R, 3)
    ((ro * go).sum() + (rd * gd).sum()).backward()
    pg = pose.detach().to(DEV).requires_grad_(True)
    oo, dd = fused_ops.camera_opt_apply(pg, frozen.to(torch.uint8).to(DEV), idx.view(-1).to(DEV), o.to(DEV), d.to(DEV),
                                        shared)
    close(oo, ro, 1e-6, 1e-6)
    close(dd, rd, 1e-6, 1e-6)
    ((oo * go.to(DEV)).sum() + (dd * gd.to(DEV)).sum()).backward()
    close(pg.grad, pose.grad, 2e-4, 1e-4)
    # and through the module, fused vs torch expression
    opt = tn.CameraOptimizer(tn.CameraOptimizerConfig(mode="SO3xR3"), cams,
                             non_trainable_camera_indices=torch.tensor([5, 6, 7])).to(DEV)
    with torch.no_grad():
        opt.pose_adjustment.normal_(0, 0.05)
    rb1 = tn.RayBundle(origins=o.to(DEV), directions=d.to(DEV), pixel_area=torch.ones(R, 1, device=DEV),
                       camera_indices=idx.to(DEV))
    rb2 = tn.RayBundle(origins=o.to(DEV), directions=d.to(DEV), pixel_area=torch.ones(R, 1, device=DEV),
                       camera_indices=idx.to(DEV))
    opt.apply_to_raybundle(rb1)
    opt.fused = False
    opt.apply_to_raybundle(rb2)
    close(rb1.origins, rb2.origins, 1e-6, 1e-6)
    close(rb1.directions, rb2.directions, 1e-6, 1e-6)


@pytest.mark.parametrize("with_thermal", [True, False])
def test_pixel_losses_kernel(with_thermal):
    torch.manual_seed(6)
    R = 512
    is_thermal = (torch.arange(R // 4) % 3 == 1).float().repeat_interleave(4)
    rgb = torch.rand(R, 3, requires_grad=True)
    th = torch.rand(R, 1, requires_grad=True)
    image = torch.rand(R, 3)
    cfg = oracle.OracleConfig(density_mode="shared" if with_thermal else "rgb_only")
    outs = {"rgb": rgb, "rgb_thermal": th}
    ref = oracle.thermal_nerfacto_losses({}, cfg, outs, image, is_thermal, training=False)
    up = torch.tensor([0.7, 1.3, 0.4, 2.0])
    keys = ["rgb_loss", "thermal_loss", "tv_pixel_loss", "cross_channel_loss"]
    mults = [1.0, cfg.thermal_loss_mult, cfg.tv_pixel_loss_mult, cfg.cross_channel_loss_mult]
    total = sum(up[i] * ref[k] / mults[i] for i, k in enumerate(keys) if k in ref)
    total.backward()
    rg = rgb.detach().to(DEV).requires_grad_(True)
    tg = th.detach().to(DEV).requires_grad_(True)
    pl = fused_ops.pixel_losses(rg, tg if with_thermal else None, image.to(DEV), is_thermal.to(DEV))
    for i, k in enumerate(keys):
        if k in ref:
            close(pl[i], ref[k] / mults[i], 1e-7, 1e-5)
    (torch.stack(pl) * up.to(DEV)).sum().backward()
    close(rg.grad, rgb.grad, 1e-8, 1e-4)
    if with_thermal:
        close(tg.grad, th.grad, 1e-8, 1e-4)


@pytest.mark.parametrize("rgb_mult", [0.01, 1.0])
def test_density_l1_kernel(rgb_mult):
    torch.manual_seed(7)
    shape = (37, 48, 1)
    ts = [(torch.rand(shape) * 5).requires_grad_(True) for _ in range(4)]
    cfg = oracle.OracleConfig(density_mode="separate", rgb_density_loss_mult=rgb_mult, tv_pixel_loss_mult=0,
                              cross_channel_loss_mult=0)
    outs = {"density": ts[0], "density2": ts[1], "density_thermal": ts[2], "density2_thermal": ts[3],
            "rgb": torch.zeros(4, 3), "rgb_thermal": torch.zeros(4, 1)}
    ref = oracle.thermal_nerfacto_losses({}, cfg, outs, torch.zeros(4, 3), torch.zeros(4), training=False)["density_loss"]
    ref.backward()
    gs = [t.detach().to(DEV).requires_grad_(True) for t in ts]
    got = fused_ops.density_l1(gs[0], gs[1], gs[2], gs[3], cfg.density_loss_mult, rgb_mult)
    close(got, ref, 1e-9, 1e-5)
    got.backward()
    for a, b in zip(gs, ts):
        close(a.grad, b.grad, 1e-12, 1e-5)


@pytest.mark.parametrize("rays,samples,out_dim", [(37, 48, 3), (11, 19, 1), (64, 256, 4)])
def test_field_head_fused_backward_equals_unfused(rays, samples, out_dim):
    """_FieldHeadFn (one backward kernel: restricted dX, density backward and per-ray dZ1 sums by a one-hot MMA) against
    the unfused composition field_split -> ops.mlp, which the goldens pin."""
    from nerfstudio_thermal_b200 import ops
    torch.manual_seed(rays)
    n, geo, emb_dim = rays * samples, 15, 32
    h = (torch.randn(n, 16, device=DEV) * 1.5).requires_grad_(True)
    sel = (torch.rand(n, device=DEV) > 0.2).float()
    sh = torch.randn(rays, 16, device=DEV)
    emb = torch.randn(rays, emb_dim, device=DEV).requires_grad_(True)
    dims = [63, 64, 64, out_dim]
    ws = [(torch.randn(dims[i + 1], dims[i], device=DEV) / dims[i] ** 0.5).requires_grad_(True) for i in range(3)]
    bs = [(torch.randn(dims[i + 1], device=DEV) * 0.1).requires_grad_(True) for i in range(3)]
    gd, gy = torch.randn(n, device=DEV), torch.randn(n, out_dim, device=DEV)
    # rays of the other modality carry an exactly zero colour gradient (the RGB loss is masked per camera): whole
    # tiles of zero rows take the head backward's no-product path, tiles that straddle the boundary the normal one
    gy.view(rays, samples, out_dim)[rays // 3: 2 * rays // 3 + 1] = 0.0

    def run(fused):
        for t in [h, emb, *ws, *bs]:
            t.grad = None
        if fused:
            dens, y = fused_ops.field_head(h, sel, sh, emb, rays, samples, geo, 0.9, ws, bs, ops.ACT_SIGMOID)
        else:
            dens, x = fused_ops.field_split(h, sel, sh, emb, rays, samples, geo, 0.9)
            y = ops.mlp(x, ws, bs, ops.ACT_SIGMOID)
        ((dens * gd).sum() + (y * gy).sum()).backward()
        return dens.detach(), y.detach(), [t.grad.clone() for t in [h, emb, *ws, *bs]]

    d0, y0, g0 = run(False)
    d1, y1, g1 = run(True)
    assert torch.equal(d0, d1) and torch.equal(y0, y1)  # same forward kernels
    names = ["h", "emb", "w1", "w2", "w3", "b1", "b2", "b3"]
    for name, a, b in zip(names, g0, g1):
        rel = ((a - b).double().norm() / (a.double().norm() + 1e-30)).item()
        assert rel <= 2e-4, (name, rel)
    # density-only gradient (no colour gradient reaches the head)
    for t in [h, emb, *ws, *bs]:
        t.grad = None
    dens, _ = fused_ops.field_head(h, sel, sh, emb, rays, samples, geo, 0.9, ws, bs, ops.ACT_SIGMOID)
    (dens * gd).sum().backward()
    ref = gd * 0.9 * sel * torch.exp(h.detach()[:, 0].clamp(-15, 15))
    close(h.grad[:, 0], ref, 1e-6, 1e-5)
    assert torch.count_nonzero(h.grad[:, 1:]) == 0


# ------------------------------------------------------------------------------- launch-count reductions (round 2)
def test_ray_gradient_chain_equals_autograd_accumulation():
    """A bundle read by several consumers (two proposal levels, the main field, a cross-field term): the chained
    pass-through outputs (`chain=True`: one gradient buffer handed back and added onto in place) give the gradients
    autograd's own accumulation gives."""
    from nerfstudio_thermal_b200 import ops
    from nerfstudio_thermal_b200.field_components import HashEncoding
    torch.manual_seed(11)
    R = 130
    o = (torch.randn(R, 3, device=DEV) * 0.3)
    d = torch.nn.functional.normalize(torch.randn(R, 3, device=DEV), dim=-1)
    bins = [torch.sort(torch.rand(R, s + 1, device=DEV) * 4 + 0.05, dim=-1).values for s in (96, 48, 48)]
    enc = HashEncoding(num_levels=5, min_res=16, max_res=128, log2_hashmap_size=12).to(DEV)
    with torch.no_grad():
        enc.hash_table.mul_(300.0)
    w1, b1 = torch.randn(16, 10, device=DEV) * 0.3, torch.randn(16, device=DEV) * 0.1
    w2, b2 = torch.randn(1, 16, device=DEV) * 0.3, torch.randn(1, device=DEV) * 0.1
    gs = [torch.randn(R * 96, device=DEV), torch.randn(R * 48, 3, device=DEV), torch.randn(R * 48, 3, device=DEV)]

    def run(chain):
        og, dg = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
        oc, dc = og * 1.0, dg * 1.0  # non-leaf, as after the pose correction
        prop = lambda a, b: fused_ops.prop_density(a, b, bins[0], enc.hash_table.detach(), w1, b1, w2, b2, enc.spec,  # noqa: E731
                                                   1.0, chain=chain)
        if chain:
            sig, oc, dc = prop(oc, dc)
            x1, _, oc, dc = ops.sample_positions(oc, dc, bins[1], chain=True)
            x2, _, oc, dc = ops.sample_positions(oc, dc, bins[2], chain=True)
        else:
            sig = prop(oc, dc)
            x1, _ = ops.sample_positions(oc, dc, bins[1])
            x2, _ = ops.sample_positions(oc, dc, bins[2])
        ((sig * gs[0]).sum() + (x1 * gs[1]).sum() + (x2 * gs[2]).sum()).backward()
        return og.grad, dg.grad

    (o0, d0), (o1, d1) = run(False), run(True)
    close(o1, o0, 1e-5, 1e-5)
    close(d1, d0, 1e-5, 1e-5)


def test_ray_features_equal_sh_of_normalised_directions_and_embedding_rows():
    from nerfstudio_thermal_b200 import ops
    from nerfstudio_thermal_b200.fields import get_normalized_directions
    torch.manual_seed(12)
    R, cams = 333, 9
    d = torch.nn.functional.normalize(torch.randn(R, 3, device=DEV), dim=-1)
    w = torch.randn(cams, 32, device=DEV)
    idx = torch.randint(0, cams, (R,), device=DEV)
    sh, emb = fused_ops.ray_features(d, w, idx)
    assert torch.equal(sh, ops.sh4(get_normalized_directions(d)))
    assert torch.equal(emb, w[idx])
    sh2, none = fused_ops.ray_features(d)
    assert torch.equal(sh2, sh) and none is None


@pytest.mark.parametrize("with_sink", [False, True])
def test_field_head_forms_the_embedding_gradient_itself(with_sink):
    """emb_weight/cam_idx path of _FieldHeadFn (tn_embed_bwd: [R,64]x[64,32] product + index_add in one launch, into
    a sink when given, persistent self-cleaning dz1 buffer) against the differentiable lookup + matmul path."""
    from nerfstudio_thermal_b200 import ops
    torch.manual_seed(13)
    rays, samples, out_dim, cams = 52, 48, 3, 7
    n = rays * samples
    h = (torch.randn(n, 16, device=DEV) * 1.5).requires_grad_(True)
    sel = (torch.rand(n, device=DEV) > 0.2).float()
    d = torch.nn.functional.normalize(torch.randn(rays, 3, device=DEV), dim=-1)
    table = torch.randn(cams, 32, device=DEV).requires_grad_(True)
    idx = (torch.arange(rays, device=DEV) // 4) % cams  # patch-ordered: runs of four rays per camera
    dims = [63, 64, 64, out_dim]
    ws = [(torch.randn(dims[i + 1], dims[i], device=DEV) / dims[i] ** 0.5) for i in range(3)]
    bs = [(torch.randn(dims[i + 1], device=DEV) * 0.1) for i in range(3)]
    gd, gy = torch.randn(n, device=DEV), torch.randn(n, out_dim, device=DEV)
    sh, emb_rows = fused_ops.ray_features(d, table, idx)
    # reference: differentiable lookup, gradient by the head's matmul + index_add_
    emb = fused_ops.embed_rows(table, idx)
    dens, y = fused_ops.field_head(h, sel, sh, emb, rays, samples, 15, 0.9, ws, bs, ops.ACT_SIGMOID)
    ((dens * gd).sum() + (y * gy).sum()).backward()
    ref_table, ref_h = table.grad.clone(), h.grad.clone()
    keep = fused_ops.LaunchScratch()
    sink = torch.zeros_like(table) if with_sink else None
    for rep in range(2):  # the second pass reuses the persistent dz1 buffer: it must have been left clean
        table.grad = h.grad = None
        dens2, y2 = fused_ops.field_head(h, sel, sh, emb_rows, rays, samples, 15, 0.9, ws, bs, ops.ACT_SIGMOID,
                                         emb_weight=table, cam_idx=idx, emb_sink=sink, scratch=keep)
        assert torch.equal(dens2, dens) and torch.equal(y2, y)
        ((dens2 * gd).sum() + (y2 * gy).sum()).backward()
        got = sink / (rep + 1) if with_sink else table.grad
        if with_sink:
            assert table.grad is None
        rel = ((got - ref_table).norm() / ref_table.norm()).item()
        assert rel <= 1e-5, rel
        assert torch.equal(h.grad, ref_h)
    buf = keep._bufs[("dz1", str(h.device), rays)]
    assert torch.count_nonzero(buf) == 0


def test_camera_kernels_accumulate_into_a_sink():
    torch.manual_seed(14)
    R, cams = 256, 8
    pose = (torch.randn(cams, 6, device=DEV) * 0.05).requires_grad_(True)
    idx = torch.randint(0, cams, (R,), device=DEV)
    o, d = torch.randn(R, 3, device=DEV), torch.nn.functional.normalize(torch.randn(R, 3, device=DEV), dim=-1)
    go, gd = torch.randn(R, 3, device=DEV), torch.randn(R, 3, device=DEV)

    def run(sink):
        pose.grad = None
        oo, dd = fused_ops.camera_opt_apply(pose, None, idx, o, d, False, sink=sink)
        reg, _, _ = fused_ops.camera_regularizer(pose, 0.01, 0.001, 10.0, sink=sink)
        ((oo * go).sum() + (dd * gd).sum() + 3.0 * reg).backward()
        return pose.grad

    ref = run(None).clone()
    sink = torch.zeros_like(pose)
    assert run(sink) is None
    close(sink, ref, 1e-6, 1e-5)


def test_density_l1_total_by_the_last_cta():
    torch.manual_seed(15)
    ts = [(torch.rand(4096 * 48, device=DEV) * 5).requires_grad_(True) for _ in range(4)]
    ref = fused_ops.density_l1(*ts, 0.001, 0.01)
    keep = fused_ops.LaunchScratch()
    for _ in range(3):  # the ticket word re-arms itself
        got = fused_ops.density_l1(*ts, 0.001, 0.01, scratch=keep)
        close(got, ref, 1e-9, 2e-6)
    assert keep._bufs[("l1", str(ts[0].device))].item() == 0
    got.backward()
    g = [t.grad.clone() for t in ts]
    for t in ts:
        t.grad = None
    (2.5 * ref).backward()
    for a, t in zip(g, ts):
        close(2.5 * a, t.grad, 1e-12, 1e-6)
