"""Pin the CPU oracle against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  Integer results bit-exact; float results identical op order -> tight."""
import pytest
import torch

import oracle
from oracle import model as om
from oracle import sampling as osamp


def close(a, b, atol=1e-6, rtol=1e-6):
    torch.testing.assert_close(a, b, atol=atol, rtol=rtol)


@pytest.mark.parametrize("tag,L,lo,hi,log2T", [("main19", 16, 16, 2048, 19), ("main21", 16, 16, 2048, 21),
                                               ("prop128", 5, 16, 128, 17), ("prop256", 5, 16, 256, 17)])
def test_hash_indices_bit_exact(golden, tag, L, lo, hi, log2T):
    g = golden("hash_indices.npz")
    scal = oracle.hash_scalings(L, lo, hi)
    assert torch.equal(scal, g[f"{tag}_scalings"])
    idx, off = oracle.hash_corner_indices(g["x"], scal, log2T)
    assert torch.equal(idx.to(torch.int32), g[f"{tag}_idx"])
    assert torch.equal(off, g[f"{tag}_offset"])


def test_probed_scalings():
    # SURVEY.md section 7 (probed from the reference): top level is 2047, not 2048
    assert oracle.hash_scalings(16, 16, 2048).tolist() == [16, 22, 30, 42, 58, 80, 111, 153, 212, 294, 406, 561, 776,
                                                            1072, 1482, 2047]
    assert oracle.hash_scalings(5, 16, 128).tolist() == [16, 26, 45, 76, 128]
    assert oracle.hash_scalings(5, 16, 256).tolist() == [16, 32, 64, 128, 256]


@pytest.mark.parametrize("name", ["hash_small.npz", "hash_small_prop.npz"])
def test_hash_encode_fwd_bwd(golden, name):
    g = golden(name)
    x = g["x"].clone().requires_grad_(True)
    table = g["table"].clone().requires_grad_(True)
    y = oracle.hash_encode(x, table, g["scalings"], int(g["log2T"]))
    assert torch.equal(y, g["y"])
    (y * g["dy"]).sum().backward()
    close(table.grad, g["dtable"], 1e-5, 1e-5)
    close(x.grad, g["dx"], 1e-4, 1e-5)


@pytest.mark.parametrize("tag,n,act", [("density", 2, None), ("head3", 3, "sigmoid"), ("head4", 3, "sigmoid"),
                                       ("head1", 3, "sigmoid"), ("prop", 2, None)])
def test_mlp(golden, tag, n, act):
    g = golden("components.npz")
    ws = [g[f"mlp_{tag}_w{i}"] for i in range(n)]
    bs = [g[f"mlp_{tag}_b{i}"] for i in range(n)]
    assert torch.equal(oracle.mlp_forward(g[f"mlp_{tag}_x"], ws, bs, act), g[f"mlp_{tag}_y"])


def test_sh_contraction_truncexp(golden):
    g = golden("components.npz")
    assert torch.equal(oracle.sh4_basis(g["sh_in"]), g["sh_out"])
    x = g["contract_in"].clone().requires_grad_(True)
    y = oracle.scene_contraction_linf(x)
    assert torch.equal(y, g["contract_out"])
    (y * g["contract_dy"]).sum().backward()
    close(x.grad, g["contract_dx"])
    t = g["truncexp_in"].clone().requires_grad_(True)
    e = oracle.trunc_exp(t)
    assert torch.equal(e, g["truncexp_out"])
    e.sum().backward()
    assert torch.equal(t.grad, g["truncexp_grad"])


def test_sh_orthonormal_known_answer():
    # restates the reference's only numeric KAT on this path, tests/utils/test_math.py:7-16
    torch.manual_seed(0)
    d = torch.nn.functional.normalize(torch.randn(1_000_000, 3), dim=-1)
    sh = oracle.sh4_basis(d)
    gram = sh.T @ sh / d.shape[0] * 4 * torch.pi
    torch.testing.assert_close(gram, torch.eye(16), rtol=0, atol=1.5e-2)


def test_positions_known_answer(golden):
    # reference golden vector tests/cameras/test_rays.py:11-31
    s = osamp.OracleSamples(origins=torch.ones(5, 3), directions=torch.tensor([[0.0, 1.0, 0.5]]).expand(5, 3),
                            camera_indices=None, starts=torch.ones(5, 1, 1) * 2, ends=torch.ones(5, 1, 1) * 3,
                            spacing_starts=None, spacing_ends=None, s_near=None, s_far=None)
    pos = oracle.sample_positions(s)[:, 0]
    torch.testing.assert_close(pos, torch.tensor([[1.0, 3.5, 2.25]]).expand(5, 3), atol=1e-6, rtol=0)
    assert torch.equal(pos, golden("sampling.npz")["kat_positions"])


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_sampling_chain(golden, mode):
    g = golden("sampling.npz")
    tr = mode == "train"
    o, d = g[f"{mode}_origins"], g[f"{mode}_directions"]
    jit = [g[f"{mode}_jit{i}"] if tr else None for i in range(3)]
    s0 = osamp.initial_samples(o, d, None, g[f"{mode}_nears"], g[f"{mode}_fars"], 256, jit[0])
    w0 = oracle.sample_weights(s0.deltas, g[f"{mode}_dens0"])
    s1 = oracle.pdf_resample(s0, w0, 96, jit[1])
    w1 = oracle.sample_weights(s1.deltas, g[f"{mode}_dens1"])
    s2 = oracle.pdf_resample(s1, w1, 48, jit[2])
    for tag, s in (("s0", s0), ("s1", s1), ("s2", s2)):
        assert torch.equal(s.starts, g[f"{mode}_{tag}_starts"])
        assert torch.equal(s.ends, g[f"{mode}_{tag}_ends"])
        assert torch.equal(s.spacing_starts.expand_as(s.starts), g[f"{mode}_{tag}_spacing_starts"])
        assert torch.equal(oracle.sample_positions(s), g[f"{mode}_{tag}_positions"])
    assert torch.equal(w0, g[f"{mode}_w0"]) and torch.equal(w1, g[f"{mode}_w1"])
    dens2 = g[f"{mode}_dens2"].clone().requires_grad_(True)
    w2 = oracle.sample_weights(s2.deltas, dens2)
    assert torch.equal(w2, g[f"{mode}_w2"])
    (w2 * g[f"{mode}_dw2"]).sum().backward()
    close(dens2.grad, g[f"{mode}_ddens2"])
    w2 = w2.detach()
    for C in (3, 1, 4):
        img = oracle.render_colour(g[f"{mode}_col{C}"], w2, "last_sample", tr)
        assert torch.equal(img, g[f"{mode}_img{C}"])
    for bg in ("black", "white", "random"):
        assert torch.equal(oracle.render_colour(g[f"{mode}_col3"], w2, bg, tr), g[f"{mode}_img3_{bg}"])
    assert torch.equal(oracle.render_accumulation(w2), g[f"{mode}_acc"])
    assert torch.equal(oracle.render_depth_median(w2, s2.starts, s2.ends), g[f"{mode}_depth_median"])
    assert torch.equal(oracle.render_depth_expected(w2, s2.starts, s2.ends), g[f"{mode}_depth_expected"])
    assert torch.equal(oracle.render_depth_median(w0, s0.starts, s0.ends), g[f"{mode}_depth_median0"])
    close(om.interlevel_loss([w0, w1, w2], [s0, s1, s2]), g[f"{mode}_interlevel"], 1e-7, 1e-6)
    close(om.distortion_loss([w0, w1, w2], [s0, s1, s2]), g[f"{mode}_distortion"], 1e-7, 1e-6)


def _cfg_for(mode):
    return oracle.OracleConfig(
        density_mode=mode, log2_hashmap_size=9,
        proposal_net_args_list=[{"hidden_dim": 16, "log2_hashmap_size": 8, "num_levels": 5, "max_res": 128},
                                {"hidden_dim": 16, "log2_hashmap_size": 8, "num_levels": 5, "max_res": 256}],
        is_thermal_cameras=(0, 0, 0, 0, 1, 1, 1, 1))


@pytest.mark.parametrize("mode", ["separate", "shared", "rgb_only"])
def test_model_eval(golden, mode):
    g = golden(f"model_{mode}.npz")
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd/")}
    with torch.no_grad():
        out = oracle.thermal_nerfacto_forward(sd, _cfg_for(mode), g["origins"], g["directions"], g["camera_indices"],
                                              training=False)
    keys = [k[5:] for k in g if k.startswith("eval/")]
    assert keys
    for k in keys:
        close(out[k], g[f"eval/{k}"], 1e-6, 1e-6)


@pytest.mark.parametrize("mode", ["separate", "shared", "rgb_only"])
def test_model_train_and_grads(golden, mode):
    g = golden(f"model_{mode}.npz")
    sd = {k[3:]: v.clone() for k, v in g.items() if k.startswith("sd/")}
    leaves = {}
    for k, v in sd.items():
        if v.dtype == torch.float32 and not k.endswith("aabb"):
            if k.endswith("encoding.hash_table"):  # alias of mlp_base.0.hash_table
                continue
            leaves[k] = v.requires_grad_(True)
    for k in list(sd):
        if k.endswith("encoding.hash_table"):
            sd[k] = sd[k.replace("encoding.hash_table", "mlp_base.0.hash_table")]
    jit = [g[f"jitter{i}"] for i in range(3)]
    jit_t = [g[f"jitter{i}"] for i in range(3, 6)] if mode == "separate" else None
    cfg = _cfg_for(mode)
    out = oracle.thermal_nerfacto_forward(sd, cfg, g["origins"], g["directions"], g["camera_indices"], training=True,
                                          jitters=jit, jitters_thermal=jit_t)
    for k in [k[6:] for k in g if k.startswith("train/")]:
        if k.startswith("weights") or k.startswith("sdist"):
            continue
        close(out[k], g[f"train/{k}"], 1e-6, 1e-6)
    for sfx in (("", "_thermal") if mode == "separate" else ("",)):
        for i in range(3):
            close(out[f"weights_list{sfx}"][i], g[f"train/weights{sfx}_{i}"])
            close(out[f"ray_samples_list{sfx}"][i].sdist(), g[f"train/sdist{sfx}_{i}"])
    if mode == "rgb_only":
        return
    losses = oracle.thermal_nerfacto_losses(sd, cfg, out, g["image"], g["is_thermal"], training=True)
    ref_keys = sorted(k[5:] for k in g if k.startswith("loss/"))
    assert sorted(losses) == ref_keys
    total = 0
    for k in ref_keys:
        close(torch.as_tensor(losses[k]), g[f"loss/{k}"], 1e-7, 1e-5)
        total = total + losses[k]
    total.backward()
    n = 0
    for k in [k[5:] for k in g if k.startswith("grad/")]:
        kk = k.replace("encoding.hash_table", "mlp_base.0.hash_table")
        assert leaves[kk].grad is not None, k
        close(leaves[kk].grad, g[f"grad/{k}"], 1e-7, 1e-4)
        n += 1
    assert n >= 10


def test_lazy_jitter_matches_reference_rng_order(golden):
    g = golden("model_separate.npz")
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd/")}
    torch.manual_seed(300)
    with torch.no_grad():
        out = oracle.thermal_nerfacto_forward(sd, _cfg_for("separate"), g["origins"], g["directions"],
                                              g["camera_indices"], training=True)
    close(out["rgb"], g["train/rgb"])
    close(out["rgb_thermal"], g["train/rgb_thermal"])


def test_interlevel_hinge_sends_exact_zeros_to_bounded_rays():
    """Premise of the proposal backward's early-out (csrc/tn_prop.cu): the interlevel loss is a hinge
    (model_components/losses.py:87-103), so a ray whose proposal histogram bounds the fine weights everywhere sends
    EXACTLY zero to its proposal densities; only rays that violate the bound carry a gradient."""
    from oracle import model as om
    from oracle import sampling as osamp
    torch.manual_seed(0)
    R, Sp, Sf = 2, 32, 12
    o, d = torch.zeros(R, 3), torch.nn.functional.normalize(torch.randn(R, 3), dim=-1)
    nears, fars = torch.full((R, 1), 0.05), torch.full((R, 1), 1000.0)
    prop = osamp.initial_samples(o, d, None, nears, fars, Sp, None)
    sigma = torch.full((R, Sp, 1), 0.3, requires_grad=True)
    w_prop = osamp.sample_weights(prop.deltas, sigma)
    fine = osamp.pdf_resample(prop, w_prop.detach(), Sf, None)
    # ray 0: fine weights far below the proposal mass they sit in (bounded); ray 1: above it (the hinge is active)
    w_fine = torch.stack([torch.full((Sf, 1), 1e-4), torch.full((Sf, 1), 0.9)])
    loss = om.interlevel_loss([w_prop, w_fine], [prop, fine])
    loss.backward()
    assert torch.count_nonzero(sigma.grad[0]) == 0
    assert torch.count_nonzero(sigma.grad[1]) > 0


def test_rgb_loss_sends_exact_zeros_to_rays_of_thermal_cameras():
    """Premise of the colour-head backward's tile skip (csrc/tn_mlp_tc.cu): the RGB loss is masked to the rays of RGB
    cameras (models/thermal_nerfacto.py:315-318), so the rendered RGB of a thermal camera's ray gets an exactly zero
    gradient -- while the thermal prediction gets one on every ray (thermal MSE on thermal rays, pixel TV and
    cross-channel terms on RGB rays)."""
    import oracle
    torch.manual_seed(1)
    R = 16
    is_thermal = (torch.arange(R) >= R // 2).float()
    rgb = torch.rand(R, 3, requires_grad=True)
    th = torch.rand(R, 1, requires_grad=True)
    cfg = oracle.OracleConfig(density_mode="shared")
    losses = oracle.thermal_nerfacto_losses({}, cfg, {"rgb": rgb, "rgb_thermal": th}, torch.rand(R, 3), is_thermal,
                                            training=False)
    sum(losses.values()).backward()
    assert torch.count_nonzero(rgb.grad[R // 2:]) == 0 and torch.count_nonzero(rgb.grad[:R // 2]) > 0
    assert torch.count_nonzero(th.grad[R // 2:]) > 0 and torch.count_nonzero(th.grad[:R // 2]) > 0
