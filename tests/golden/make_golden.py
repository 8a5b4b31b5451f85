"""Generate the golden fixtures in tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Everything written here is an input or an output of reference code (nerfstudio 1.0.2 fork,
`implementation="torch"`, CPU float32).  The fixtures pin the oracle (tests/test_oracle_golden.py) and,
through it and directly, the CUDA path (tests/test_gpu_*.py).  Nothing at test / bench time imports the
reference.
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_shim  # noqa: E402

_ref_shim.install()

from nerfstudio.cameras.rays import Frustums, RayBundle, RaySamples  # noqa: E402
from nerfstudio.data.scene_box import SceneBox  # noqa: E402
from nerfstudio.field_components.activations import trunc_exp  # noqa: E402
from nerfstudio.field_components.encodings import HashEncoding, SHEncoding  # noqa: E402
from nerfstudio.field_components.mlp import MLP  # noqa: E402
from nerfstudio.field_components.spatial_distortions import SceneContraction  # noqa: E402
from nerfstudio.model_components import losses as ref_losses  # noqa: E402
from nerfstudio.model_components.ray_samplers import PDFSampler, UniformLinDispPiecewiseSampler  # noqa: E402
from nerfstudio.model_components.renderers import (  # noqa: E402
    AccumulationRenderer,
    DepthRenderer,
    RGBRenderer,
    RGBTRenderer,
)
from nerfstudio.models.thermal_nerfacto import ThermalNerfactoModelConfig  # noqa: E402


def npy(t):
    return t.detach().cpu().numpy()


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: (npy(v) if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()})
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB")


def ref_hash_indices(enc: HashEncoding, x):
    """The eight hashed corner rows in the order of encodings.py:431-438, via the reference's hash_fn."""
    s = x[..., None, :] * enc.scalings.view(-1, 1)
    c = torch.ceil(s).type(torch.int32)
    f = torch.floor(s).type(torch.int32)
    pick = [(c, c, c), (c, f, c), (f, f, c), (f, c, c), (c, c, f), (c, f, f), (f, f, f), (f, c, f)]
    out = [enc.hash_fn(torch.cat([a[..., 0:1], b[..., 1:2], d[..., 2:3]], dim=-1)) for a, b, d in pick]
    return torch.stack(out, dim=-1), s - f


def edge_points():
    g = torch.tensor([0.0, 1.0, 0.5, 0.25, 1.0 / 3.0, 1e-7, 1 - 1e-7, 1 / 16, 15 / 16, 1 / 2047, 2046 / 2047])
    return torch.cartesian_prod(g, g[:4], g[2:6]).float()


# --------------------------------------------------------------------------- hash grid
def gen_hash():
    torch.manual_seed(0)
    x = torch.cat([torch.rand(192, 3), edge_points()[:64]], 0)
    out = {"x": x}
    for tag, kw in {
        "main19": dict(num_levels=16, min_res=16, max_res=2048, log2_hashmap_size=19),
        "main21": dict(num_levels=16, min_res=16, max_res=2048, log2_hashmap_size=21),
        "prop128": dict(num_levels=5, min_res=16, max_res=128, log2_hashmap_size=17),
        "prop256": dict(num_levels=5, min_res=16, max_res=256, log2_hashmap_size=17),
    }.items():
        # indices do not depend on table values: build a 1-feature table to keep memory small
        enc = HashEncoding(implementation="torch", features_per_level=1, **kw)
        idx, off = ref_hash_indices(enc, x)
        out[f"{tag}_scalings"] = enc.scalings
        out[f"{tag}_idx"] = idx.to(torch.int32)
        out[f"{tag}_offset"] = off
    save("hash_indices.npz", **out)

    # small table: full forward + backward
    torch.manual_seed(1)
    enc = HashEncoding(num_levels=16, min_res=16, max_res=2048, log2_hashmap_size=9, implementation="torch")
    with torch.no_grad():
        enc.hash_table.mul_(400.0)
    xs = torch.cat([torch.rand(160, 3), edge_points()[:32]], 0).requires_grad_(True)
    y = enc(xs)
    g = torch.randn_like(y)
    (y * g).sum().backward()
    idx, _ = ref_hash_indices(enc, xs.detach())
    save("hash_small.npz", x=xs, table=enc.hash_table, scalings=enc.scalings, log2T=9, y=y, dy=g,
         dtable=enc.hash_table.grad, dx=xs.grad, idx=idx.to(torch.int32))

    torch.manual_seed(2)
    enc = HashEncoding(num_levels=5, min_res=16, max_res=128, log2_hashmap_size=8, implementation="torch")
    with torch.no_grad():
        enc.hash_table.mul_(400.0)
    xs = torch.rand(128, 3).requires_grad_(True)
    y = enc(xs)
    g = torch.randn_like(y)
    (y * g).sum().backward()
    save("hash_small_prop.npz", x=xs, table=enc.hash_table, scalings=enc.scalings, log2T=8, y=y, dy=g,
         dtable=enc.hash_table.grad, dx=xs.grad)


# --------------------------------------------------------------------------- small components
def gen_components():
    out = {}
    torch.manual_seed(3)
    for tag, (i, n, w, o, act) in {
        "density": (32, 2, 64, 16, None),
        "head3": (63, 3, 64, 3, torch.nn.Sigmoid()),
        "head4": (63, 3, 64, 4, torch.nn.Sigmoid()),
        "head1": (63, 3, 64, 1, torch.nn.Sigmoid()),
        "prop": (10, 2, 16, 1, None),
    }.items():
        mlp = MLP(in_dim=i, num_layers=n, layer_width=w, out_dim=o, out_activation=act, implementation="torch")
        x = torch.randn(96, i).requires_grad_(True)
        y = mlp(x)
        g = torch.randn_like(y)
        (y * g).sum().backward()
        out[f"mlp_{tag}_x"], out[f"mlp_{tag}_y"], out[f"mlp_{tag}_dy"], out[f"mlp_{tag}_dx"] = x, y, g, x.grad
        for li, layer in enumerate(mlp.layers):
            out[f"mlp_{tag}_w{li}"], out[f"mlp_{tag}_b{li}"] = layer.weight, layer.bias
            out[f"mlp_{tag}_dw{li}"], out[f"mlp_{tag}_db{li}"] = layer.weight.grad, layer.bias.grad

    d = torch.nn.functional.normalize(torch.randn(128, 3), dim=-1)
    out["sh_in"] = (d + 1) / 2
    out["sh_out"] = SHEncoding(levels=4, implementation="torch")(out["sh_in"])

    p = torch.randn(256, 3) * 2.0
    p[:16] *= 50.0
    p[16:24] = torch.tensor([1.0, -1.0, 0.5])
    p.requires_grad_(True)
    c = SceneContraction(order=float("inf"))(p)
    gc = torch.randn_like(c)
    (c * gc).sum().backward()
    out["contract_in"], out["contract_out"], out["contract_dy"], out["contract_dx"] = p, c, gc, p.grad

    t = (torch.randn(128) * 8).requires_grad_(True)
    e = trunc_exp(t)
    e.sum().backward()
    out["truncexp_in"], out["truncexp_out"], out["truncexp_grad"] = t, e, t.grad
    save("components.npz", **out)


# --------------------------------------------------------------------------- samplers / weights / renderers
def make_bundle(R, seed, near, far=1000.0):
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(R, 3, generator=g) * 0.3
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    cams = torch.arange(R)[:, None] % 8
    return RayBundle(origins=o, directions=d, pixel_area=torch.full((R, 1), 1e-6), camera_indices=cams,
                     nears=torch.full((R, 1), near), fars=torch.full((R, 1), far))


def samples_dict(tag, rs: RaySamples):
    return {
        f"{tag}_starts": rs.frustums.starts, f"{tag}_ends": rs.frustums.ends, f"{tag}_deltas": rs.deltas,
        f"{tag}_spacing_starts": rs.spacing_starts, f"{tag}_spacing_ends": rs.spacing_ends,
        f"{tag}_positions": rs.frustums.get_positions(),
    }


def gen_sampling():
    out = {}
    R = 24
    for mode in ("train", "eval"):
        training = mode == "train"
        rb = make_bundle(R, 10, 0.05 if training else 0.0)
        out[f"{mode}_origins"], out[f"{mode}_directions"] = rb.origins, rb.directions
        out[f"{mode}_nears"], out[f"{mode}_fars"] = rb.nears, rb.fars
        init = UniformLinDispPiecewiseSampler(single_jitter=True).train(training)
        pdf = PDFSampler(include_original=False, single_jitter=True).train(training)
        torch.manual_seed(77)
        s0 = init(rb, num_samples=256)
        torch.manual_seed(78)
        dens0 = torch.rand(R, 256, 1) ** 4 * 30.0
        dens0[:2] = 0.0  # zero-density rays
        dens0[2:4, 100:] = 1e4  # opaque wall
        w0 = s0.get_weights(dens0)
        torch.manual_seed(79)
        s1 = pdf(rb, s0, w0, num_samples=96)
        torch.manual_seed(80)
        dens1 = torch.rand(R, 96, 1) ** 2 * 50.0
        w1 = s1.get_weights(dens1)
        torch.manual_seed(81)
        s2 = pdf(rb, s1, w1, num_samples=48)
        torch.manual_seed(77)
        out[f"{mode}_jit0"] = torch.rand(R, 1)
        torch.manual_seed(79)
        out[f"{mode}_jit1"] = torch.rand(R, 1)
        torch.manual_seed(81)
        out[f"{mode}_jit2"] = torch.rand(R, 1)
        out.update(samples_dict(f"{mode}_s0", s0))
        out.update(samples_dict(f"{mode}_s1", s1))
        out.update(samples_dict(f"{mode}_s2", s2))
        out[f"{mode}_dens0"], out[f"{mode}_w0"] = dens0, w0
        out[f"{mode}_dens1"], out[f"{mode}_w1"] = dens1, w1

        # weights backward + renderers on the last level
        torch.manual_seed(82)
        dens2 = (torch.rand(R, 48, 1) ** 2 * 80.0).requires_grad_(True)
        w2 = s2.get_weights(dens2)
        gw = torch.randn_like(w2)
        (w2 * gw).sum().backward()
        out[f"{mode}_dens2"], out[f"{mode}_w2"], out[f"{mode}_dw2"], out[f"{mode}_ddens2"] = dens2, w2, gw, dens2.grad
        w2 = w2.detach()
        for C, rend in ((3, RGBRenderer(background_color="last_sample")),
                        (1, RGBRenderer(background_color="last_sample", num_channels=1)),
                        (4, RGBTRenderer(background_color="last_sample"))):
            rend.train(training)
            col = torch.rand(R, 48, C).requires_grad_(True)
            wv = w2.clone().requires_grad_(True)
            img = rend(rgb=col, weights=wv)
            gi = torch.randn_like(img)
            (img * gi).sum().backward()
            out[f"{mode}_col{C}"], out[f"{mode}_img{C}"], out[f"{mode}_dimg{C}"] = col, img, gi
            out[f"{mode}_dcol{C}"], out[f"{mode}_dw_from_img{C}"] = col.grad, wv.grad
        for bg in ("black", "white", "random"):
            rend = RGBRenderer(background_color=bg).train(training)
            out[f"{mode}_img3_{bg}"] = rend(rgb=out[f"{mode}_col3"].detach(), weights=w2)
        out[f"{mode}_acc"] = AccumulationRenderer()(weights=w2)
        out[f"{mode}_depth_median"] = DepthRenderer("median")(weights=w2, ray_samples=s2)
        out[f"{mode}_depth_expected"] = DepthRenderer("expected")(weights=w2, ray_samples=s2)
        out[f"{mode}_depth_median0"] = DepthRenderer("median")(weights=w0, ray_samples=s0)

        # per-ray losses on these levels
        wl = [w0.clone().requires_grad_(True), w1.clone().requires_grad_(True), w2.clone().requires_grad_(True)]
        il = ref_losses.interlevel_loss(wl, [s0, s1, s2])
        dl = ref_losses.distortion_loss(wl, [s0, s1, s2])
        (il + dl).backward()
        out[f"{mode}_interlevel"], out[f"{mode}_distortion"] = il, dl
        out[f"{mode}_dinter_w0"], out[f"{mode}_dinter_w1"], out[f"{mode}_ddist_w2"] = wl[0].grad, wl[1].grad, wl[2].grad
    # reference known-answer test tests/cameras/test_rays.py:11-31
    fr = Frustums(origins=torch.ones((5, 3)), directions=torch.tensor([[0.0, 1.0, 0.5]]).expand(5, 3),
                  starts=torch.ones((5, 1)) * 2, ends=torch.ones((5, 1)) * 3, pixel_area=torch.ones((5, 1)))
    out["kat_positions"] = fr.get_positions()
    save("sampling.npz", **out)


# --------------------------------------------------------------------------- full model
def make_batch(R, num_cams, seed):
    """Patch-structured batch: groups of 4 rays share one camera; RGB cameras first."""
    g = torch.Generator().manual_seed(seed)
    P = R // 4
    cam_of_patch = torch.arange(P) * num_cams // P
    cams = cam_of_patch.repeat_interleave(4)[:, None]
    centre = torch.randn(P, 3, generator=g) * 0.3
    o = centre.repeat_interleave(4, 0) + torch.randn(R, 3, generator=g) * 0.01
    dirs = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=-1).repeat_interleave(4, 0)
    d = torch.nn.functional.normalize(dirs + torch.randn(R, 3, generator=g) * 0.01, dim=-1)
    image = torch.rand(R, 3, generator=g)
    is_thermal = (cams[:, 0] >= num_cams // 2).float()
    return o, d, cams, image, is_thermal


def gen_model(mode, R=32, num_cams=8):
    torch.manual_seed(100)
    cfg = ThermalNerfactoModelConfig(
        implementation="torch", density_mode=mode, log2_hashmap_size=9,
        proposal_net_args_list=[
            {"hidden_dim": 16, "log2_hashmap_size": 8, "num_levels": 5, "max_res": 128, "use_linear": False},
            {"hidden_dim": 16, "log2_hashmap_size": 8, "num_levels": 5, "max_res": 256, "use_linear": False},
        ],
    )
    is_th = [0] * (num_cams // 2) + [1] * (num_cams // 2)
    model = cfg.setup(scene_box=SceneBox(aabb=torch.tensor([[-1.0, -1, -1], [1, 1, 1]])), num_train_data=num_cams,
                      metadata={"is_thermal": is_th})
    with torch.no_grad():  # "trained-like" weights so that weights/depths are non-degenerate
        for k, p in model.named_parameters():
            if k.endswith("hash_table"):
                p.mul_(400.0)
            if k.endswith("pose_adjustment"):
                p.normal_(0.0, 0.01)
    out = {f"sd/{k}": v for k, v in model.state_dict().items() if k != "device_indicator_param"}
    o, d, cams, image, is_thermal = make_batch(R, num_cams, 200)
    out.update(origins=o, directions=d, camera_indices=cams, image=image, is_thermal=is_thermal)

    def bundle():
        return RayBundle(origins=o.clone(), directions=d.clone(), pixel_area=torch.full((R, 1), 1e-6),
                         camera_indices=cams.clone())

    # ---- eval
    model.eval()
    with torch.no_grad():
        ev = model(bundle())
    for k, v in ev.items():
        if torch.is_tensor(v):
            out[f"eval/{k}"] = v
    # ---- train
    model.train()
    torch.manual_seed(300)
    tr = model(bundle())
    torch.manual_seed(300)
    n_draws = 6 if mode == "separate" else 3
    for i in range(n_draws):
        out[f"jitter{i}"] = torch.rand(R, 1)
    for k, v in tr.items():
        if torch.is_tensor(v):
            out[f"train/{k}"] = v
    for sfx in (("", "_thermal") if mode == "separate" else ("",)):
        for i, (w, rs) in enumerate(zip(tr[f"weights_list{sfx}"], tr[f"ray_samples_list{sfx}"])):
            out[f"train/weights{sfx}_{i}"] = w
            out[f"train/sdist{sfx}_{i}"] = ref_losses.ray_samples_to_sdist(rs)
    if mode != "rgb_only":  # rgb_only's get_loss_dict hard-codes .to("cuda") (thermal_nerfacto.py:297)
        batch = {"image": image, "is_thermal": is_thermal}
        md = model.get_metrics_dict(tr, batch)
        ld = model.get_loss_dict(tr, batch, md)
        total = 0
        for k, v in ld.items():
            out[f"loss/{k}"] = torch.as_tensor(v)
            total = total + v
        total.backward()
        seen = set()
        for k, p in model.named_parameters():
            if p.grad is not None and id(p) not in seen:
                seen.add(id(p))
                out[f"grad/{k}"] = p.grad
    save(f"model_{mode}.npz", **out)


if __name__ == "__main__":
    gen_hash()
    gen_components()
    gen_sampling()
    for m in ("separate", "shared", "rgb_only"):
        gen_model(m)
