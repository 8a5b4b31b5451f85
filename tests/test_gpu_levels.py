"""GPU tests of the one-launch-per-level kernels (csrc/tn_level.cu).

The single-purpose kernels (tn_weights_*, tn_render_*, tn_pdf_sample, tn_distortion_loss, tn_interlevel_loss) are
checked against the reference's golden vectors and the oracle in test_gpu_kernels.py / test_gpu_fused.py; the fused
launches perform the same operations in the same order, so they must reproduce those kernels EXACTLY (forward) and
to rounding (backward, where sums are formed in a different order).  The model-level test checks the fused branch
against the component-by-component branch and against the oracle on the same weights, rays and jitter."""
import pytest
import torch

import oracle
import oracle.sampling

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import nerfstudio_thermal_b200 as tn
    from nerfstudio_thermal_b200 import fused_ops, ops

DEV = "cuda"


def _level(R, S, seed, opaque=False):
    g = torch.Generator().manual_seed(seed)
    sb = torch.sort(torch.rand(R, S + 1, generator=g), dim=-1).values
    sb[:, 0], sb[:, -1] = 0.0, 1.0
    nears, fars = torch.full((R,), 0.05), torch.full((R,), 1000.0)
    sn, sf = oracle.sampling.piecewise_spacing(nears)[:, None], oracle.sampling.piecewise_spacing(fars)[:, None]
    eb = oracle.sampling.piecewise_spacing_inv(sb * sf + (1 - sb) * sn)
    sigma = torch.rand(R, S, generator=g) * (50.0 if opaque else 3.0)
    sigma[0] = 0.0  # an empty ray
    return sigma, eb.contiguous(), sb.contiguous(), nears, fars


@pytest.mark.parametrize("S,S_new,train", [(256, 96, True), (96, 48, True), (100, 37, False), (31, 48, True)])
def test_level_resample_equals_the_separate_kernels(S, S_new, train):
    R = 301
    sigma, eb, sb, nears, fars = (t.to(DEV) for t in _level(R, S, 3 + S))
    jitter = torch.rand(R, 1, device=DEV) if train else None
    anneal = torch.full((1,), 0.37, device=DEV)
    w_ref = ops.sample_weights(sigma, eb[:, 1:] - eb[:, :-1])
    _, _, med_ref, _, _ = ops.render(w_ref, None, None, None, want_depth=True, bins=eb)
    sb_ref, eb_ref = ops.pdf_sample(w_ref, sb, nears, fars, S_new, jitter, anneal=anneal)
    w, med, sb_new, eb_new = fused_ops.level_resample(sigma, eb, sb, nears, fars, S_new, jitter, anneal=anneal)
    assert torch.equal(w, w_ref) and torch.equal(med, med_ref)
    assert torch.equal(sb_new, sb_ref) and torch.equal(eb_new, eb_ref)


BG_NONE, BG_LAST_SAMPLE, BG_CONSTANT = 0, 1, 2  # ops.BG_* (the package imports only where a GPU is present)


@pytest.mark.parametrize("C,bg_mode,bg,eval_mode", [(3, BG_LAST_SAMPLE, None, False), (1, BG_LAST_SAMPLE, None, False),
                                                   (4, BG_CONSTANT, (1.0, 1.0, 1.0, 1.0), False),
                                                   (3, BG_NONE, None, True)])
def test_ray_heads_equal_the_separate_kernels(C, bg_mode, bg, eval_mode):
    from nerfstudio_thermal_b200 import fused_ops as fo

    R, S, props = 203, 48, (256, 96)
    sigma, eb, sb, _, _ = (t.to(DEV) for t in _level(R, S, 11, opaque=True))
    colour = torch.rand(R, S, C, device=DEV)
    plev = [tuple(t.to(DEV) for t in _level(R, sp, 20 + sp)) for sp in props]
    # reference composition out of the single-purpose ops, each an autograd node
    sig_r, col_r = sigma.clone().requires_grad_(True), colour.clone().requires_grad_(True)
    psig_r = [p[0].clone().requires_grad_(True) for p in plev]
    w_r = ops.sample_weights(sig_r, eb[:, 1:] - eb[:, :-1])
    rgb_r, acc_r, med_r, exp_r, mm_r = ops.render(w_r, col_r, None, None, bg_mode=bg_mode, bg=bg, eval_mode=eval_mode,
                                                  want_depth=True, bins=eb)
    dist_r = fo.distortion_loss_rays(w_r, sb)
    inter_r = 0.0
    pw = []
    for (ps, peb, psb, _, _), pg in zip(plev, psig_r):
        wp = ops.sample_weights(pg, peb[:, 1:] - peb[:, :-1])
        pw.append(wp.detach())
        inter_r = inter_r + fo.interlevel_loss_level(w_r, sb, wp, psb)
    exp_r = torch.clamp(exp_r, mm_r[0], mm_r[1])  # ray_heads returns the clipped depth (renderers.py:574)
    g_rgb, g_acc, g_exp = torch.randn_like(rgb_r), torch.randn_like(acc_r), torch.randn_like(exp_r)
    loss_r = (rgb_r * g_rgb).sum() + (acc_r * g_acc).sum() + (exp_r * g_exp).sum() + 0.3 * dist_r + 1.7 * inter_r
    loss_r.backward()

    # both launch protocols: self-initialising (clone before, scale + clamp after the kernel) and with the call
    # site's persistent scratch words, which the kernel's last CTA re-arms -- the SECOND and THIRD pass reuse them
    keep = fo.LaunchScratch()
    for scratch in (None, fo.heads_scratch(keep, DEV, 0), fo.heads_scratch(keep, DEV, 0)):
        sig, col = sigma.clone().requires_grad_(True), colour.clone().requires_grad_(True)
        psig = [p[0].clone().requires_grad_(True) for p in plev]
        rgb, acc, med, exp, mm, w, dist, inter = fo.ray_heads(
            sig, col, eb, sb, bg_mode=bg_mode, bg=bg, eval_mode=eval_mode, want_losses=True, prop_sigma=psig,
            prop_ebins=[p[1] for p in plev], prop_sbins=[p[2] for p in plev], prop_weights=pw, scratch=scratch)
        assert torch.equal(w, w_r.detach()) and torch.equal(rgb, rgb_r.detach()) and torch.equal(acc, acc_r.detach())
        assert torch.equal(med, med_r) and torch.equal(exp, exp_r.detach()) and torch.equal(mm, mm_r)
        torch.testing.assert_close(dist, dist_r.detach(), rtol=2e-6, atol=1e-9)  # sums over rays in another order
        torch.testing.assert_close(inter, inter_r.detach(), rtol=2e-6, atol=1e-9)
        ((rgb * g_rgb).sum() + (acc * g_acc).sum() + (exp * g_exp).sum() + 0.3 * dist + 1.7 * inter).backward()
        torch.testing.assert_close(sig.grad, sig_r.grad, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(col.grad, col_r.grad, rtol=1e-6, atol=1e-7)
        for a, b in zip(psig, psig_r):
            torch.testing.assert_close(a.grad, b.grad, rtol=1e-5, atol=1e-8)
        if scratch is not None:  # left in its initial state
            assert scratch[:4].tolist() == [float("inf"), float("-inf"), 0.0, 0.0]
            assert scratch.view(torch.int32)[4].item() == 0


def test_ray_heads_without_losses_or_proposal_gradients():
    from nerfstudio_thermal_b200 import fused_ops as fo

    R, S = 77, 48
    sigma, eb, sb, _, _ = (t.to(DEV) for t in _level(R, S, 5))
    colour = torch.rand(R, S, 3, device=DEV)
    with torch.no_grad():  # eval: no losses
        out = fo.ray_heads(sigma, colour, eb, sb, bg_mode=BG_LAST_SAMPLE, bg=None, eval_mode=True, want_losses=False)
    assert out[6] is None and out[7] is None and out[0].shape == (R, 3)
    # training while the proposal networks are not updated: the interlevel VALUE is still formed, no gradient flows
    psig, peb, psb, _, _ = (t.to(DEV) for t in _level(R, 96, 6))
    pw = ops.sample_weights(psig, peb[:, 1:] - peb[:, :-1])
    sig = sigma.clone().requires_grad_(True)
    out = fo.ray_heads(sig, colour, eb, sb, bg_mode=BG_LAST_SAMPLE, bg=None, eval_mode=False, want_losses=True,
                       prop_sigma=(), prop_ebins=[peb], prop_sbins=[psb], prop_weights=[pw])
    ref = fo.interlevel_loss_level(out[5], sb, pw, psb)
    torch.testing.assert_close(out[7], ref, rtol=2e-6, atol=1e-9)
    (out[0].sum() + out[7]).backward()
    assert sig.grad is not None and torch.isfinite(sig.grad).all()


def _small_model(mode, seed=0):
    torch.manual_seed(seed)
    props = [{"hidden_dim": 16, "log2_hashmap_size": 10, "num_levels": 5, "max_res": m, "use_linear": False}
             for m in (128, 256)]
    cfg = tn.ThermalNerfactoModelConfig(density_mode=mode, log2_hashmap_size=12, proposal_net_args_list=props)
    model = cfg.setup(num_train_data=8, metadata={"is_thermal": [0] * 4 + [1] * 4})
    with torch.no_grad():
        for k, p in model.named_parameters():
            if k.endswith("hash_table"):
                p.mul_(300.0)
            if "pose_adjustment" in k:
                p.normal_(0, 1e-3)
    return model.to(DEV)


@pytest.mark.parametrize("mode", ["separate", "shared", "rgb_only"])
@pytest.mark.parametrize("updated", [True, False])
def test_fused_branch_equals_component_branch(mode, updated):
    """ThermalNerfactoModel with per-level launches (fuse_levels) vs the component-by-component path: same outputs,
    same loss terms, same parameter gradients; also while the proposal networks are not being updated."""
    model = _small_model(mode).train()
    R = 64
    g = torch.Generator().manual_seed(1)
    cams = (torch.arange(R // 4) * 8 // (R // 4)).repeat_interleave(4)[:, None]
    o = torch.randn(R, 3, generator=g) * 0.3
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    batch = {"image": torch.rand(R, 3, generator=g).to(DEV), "is_thermal": (cams[:, 0] >= 4).float().to(DEV)}
    jit = [torch.rand(R, 1, generator=g).to(DEV) for _ in range(6)]
    res = {}
    for fused in (False, True):
        model.fuse_levels = fused
        model.zero_grad(set_to_none=True)
        for smp in (model.proposal_sampler, model.proposal_sampler_thermal):
            smp._forced_updated = updated
            smp.set_anneal(0.6)
        rb = tn.RayBundle(origins=o.to(DEV), directions=d.to(DEV), pixel_area=torch.ones(R, 1, device=DEV),
                          camera_indices=cams.to(DEV))
        out, losses, _ = model.get_train_loss_dict(rb, batch, jitters=jit[:3], jitters_thermal=jit[3:])
        sum(losses.values()).backward()
        res[fused] = (out, {k: float(v) for k, v in losses.items()},
                      {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
    out_a, loss_a, grad_a = res[False]
    out_b, loss_b, grad_b = res[True]
    for k, v in out_a.items():
        if torch.is_tensor(v) and not k.startswith("_"):
            assert torch.equal(v, out_b[k]), k
    for k in ("weights_list", "weights_list_thermal"):
        if k in out_a:
            for a, b in zip(out_a[k], out_b[k]):
                assert torch.equal(a, b), k
    assert loss_a.keys() == loss_b.keys()
    for k in loss_a:
        assert abs(loss_a[k] - loss_b[k]) <= 2e-6 * max(abs(loss_a[k]), 1e-6), (k, loss_a[k], loss_b[k])
    assert grad_a.keys() == grad_b.keys()
    for k in grad_a:
        err = (grad_a[k] - grad_b[k]).norm() / (grad_a[k].norm() + 1e-30)
        assert err <= 2e-5, (k, float(err))
    if not updated:  # the proposal networks received no gradient at all (engine/optimizers.py skips them)
        assert not any("proposal_networks" in k for k in grad_b)


def test_fused_branch_eval_outputs_equal_component_branch():
    model = _small_model("separate").eval()
    R = 96
    g = torch.Generator().manual_seed(2)
    rb = dict(origins=(torch.randn(R, 3, generator=g) * 0.3).to(DEV),
              directions=torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(DEV),
              pixel_area=torch.ones(R, 1, device=DEV), camera_indices=torch.zeros(R, 1, dtype=torch.long, device=DEV))
    outs = []
    for fused in (False, True):
        model.fuse_levels = fused
        with torch.no_grad():
            outs.append(model(tn.RayBundle(**rb)))
    assert {k for k in outs[0] if not k.startswith("_")} == {k for k in outs[1] if not k.startswith("_")}
    for k, v in outs[0].items():
        if torch.is_tensor(v) and not k.startswith("_"):
            assert torch.equal(v, outs[1][k]), k
