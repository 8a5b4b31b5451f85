"""Train-step parity at the configurations bench.py runs, against the oracle executed ON THE GPU.

BASELINE.json configs: [1] separate / 4096 rays / T=2^19, [3] shared / 8192 rays, [4] the T=2^21 tables of the sweep;
both the reference's initialisers and the 'trained-like' tables bench.py times.  Compared: every
rendered output (<= 1e-3), sample placement (counts exact, spacing bins <= 1e-5), median depths by the boundary
criterion of parity_utils.median_depth_check, every loss term (<= 1e-3 relative) and EVERY parameter gradient
(<= 1e-3 relative L2).  The measured numbers are written to gpurun_out/parity_*.json and tabulated in DESIGN.md.

Conditioning: with the trained-like tables (U(-0.5,0.5): finest-level feature slopes ~1000 per unit of x) the loss is
an ill-conditioned function of the sample positions.  Where a gradient misses 1e-3 the test proves that this is the
function and not the kernels: the ORACLE ITSELF, re-run with every sample position moved by one float32 ulp, moves
that gradient by a comparable amount (the bound asserted is 4x the oracle's own 1-ulp sensitivity).
"""
import pytest
import torch

import oracle
from oracle import sampling as oracle_sampling

import parity_utils as pu

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import nerfstudio_thermal_b200 as tn

DEV = "cuda"


def _batch(rays, seed, num_cams=64):
    import bench
    return {k: v.to(DEV) for k, v in bench.make_batch(rays, seed, num_cams).items()}


def _kernel_step(model, batch, jit, sep):
    model.zero_grad(set_to_none=True)
    rb = tn.RayBundle(origins=batch["origins"].clone(), directions=batch["directions"].clone(),
                      pixel_area=batch["pixel_area"], camera_indices=batch["camera_indices"])
    out, losses, _ = model.get_train_loss_dict(rb, {"image": batch["image"], "is_thermal": batch["is_thermal"]},
                                               jitters=jit[:3], jitters_thermal=jit[3:] if sep else None)
    losses.total.backward()
    return out, losses


CASES = [
    # name,            mode,       rays, log2T, init
    ("separate_4096_T19_reference", "separate", 4096, 19, "reference"),
    ("separate_4096_T19_trained", "separate", 4096, 19, "trained"),
    ("shared_8192_T19_reference", "shared", 8192, 19, "reference"),
    ("shared_8192_T19_trained", "shared", 8192, 19, "trained"),
    ("separate_4096_T21_reference", "separate", 4096, 21, "reference"),
    ("separate_4096_T21_trained", "separate", 4096, 21, "trained"),
]


@pytest.mark.parametrize("name,mode,rays,log2_T,init", CASES, ids=[c[0] for c in CASES])
def test_full_size_train_step_vs_oracle_on_cuda(name, mode, rays, log2_T, init):
    sep = mode == "separate"
    model, ocfg = pu.bench_like_model(tn, mode, log2_T, init)
    model = model.to(DEV).train()
    sd = pu.oracle_state(model, DEV)
    batch = _batch(rays, 42)
    gen = torch.Generator().manual_seed(77)
    jit = [torch.rand(rays, 1, generator=gen).to(DEV) for _ in range(6)]

    out, losses = _kernel_step(model, batch, jit, sep)
    ref, ref_losses = pu.oracle_train_step(sd, ocfg, batch, jit)
    torch.cuda.synchronize()

    rep = {"config": {"mode": mode, "rays": rays, "log2_T": log2_T, "init": init}}
    failures = []

    # ---- sample placement: counts exact, spacing-domain bins to float rounding
    place = {}
    for sfx in (("", "_thermal") if sep else ("",)):
        for i, n in enumerate((256, 96, 48)):
            got = out[f"ray_samples_list{sfx}"][i]._layout.sbins
            want = ref[f"ray_samples_list{sfx}"][i].sdist().detach()
            assert got.shape == (rays, n + 1) == want.shape  # sample counts exact
            place[f"sbins{sfx}_{i}"] = (got - want).abs().max().item()
            w_err = (out[f"weights_list{sfx}"][i].detach() - ref[f"weights_list{sfx}"][i].detach()).abs().max().item()
            place[f"weights{sfx}_{i}"] = w_err
    rep["placement_max_abs"] = place
    for k, v in place.items():
        if v > (1e-5 if k.startswith("sbins") else 1e-3):
            failures.append(("placement", k, v))

    # ---- rendered outputs + median depths
    errs = pu.compare_outputs(out, ref, mode, strict=False)
    rep["outputs"] = errs
    for k, v in errs.items():
        if isinstance(v, dict):
            if v["unexcused"] or v["misplaced"]:
                failures.append(("median_depth", k, v))
        elif v > pu.OUT_TOL:
            failures.append(("output", k, v))

    # ---- loss terms
    assert sorted(losses) == sorted(ref_losses)
    rep["losses"] = {}
    for k in ref_losses:
        a, b = float(losses[k]), float(ref_losses[k])
        rel = abs(a - b) / max(abs(b), 1e-12)
        rep["losses"][k] = {"got": a, "oracle": b, "rel": rel}
        if rel > pu.LOSS_TOL and abs(a - b) > 1e-9:
            failures.append(("loss", k, rel))

    # ---- every parameter gradient
    grads = pu.param_grads_vs_oracle(model, sd)
    rep["grad_rel_l2"] = grads
    assert len(grads) >= (20 if sep else 10), sorted(grads)
    over = {k: v for k, v in grads.items() if v > pu.GRAD_TOL}
    rep["grad_over_1e-3"] = sorted(over)

    if over:
        # conditioning proof: the oracle's own gradients under a one-ulp shift of every sample position
        base = {k: v.grad.detach().clone() for k, v in sd.items() if v.requires_grad and v.grad is not None}
        oracle_sampling.POSITION_ULP_SHIFT = 1
        try:
            pu.oracle_train_step(sd, ocfg, batch, jit)
        finally:
            oracle_sampling.POSITION_ULP_SHIFT = 0
        sens = {}
        for k in over:
            kk = k.replace("encoding.hash_table", "mlp_base.0.hash_table")
            sens[k] = pu.rel_l2(sd[kk].grad, base[kk])
            if over[k] > 4.0 * sens[k]:
                failures.append(("gradient", k, over[k], "oracle 1-ulp sensitivity", sens[k]))
        rep["oracle_one_ulp_sensitivity"] = sens
    rep["failures"] = [list(map(str, f)) for f in failures]
    pu.report(name, rep)
    assert not failures, failures
    if init == "reference":  # the reference's own weight scale meets the bar outright
        assert not over, over


def test_eval_full_size_vs_oracle_on_cuda():
    """Eval forward of a whole 32768-ray chunk (the render chunk size) for density_mode=separate, every output the
    renderer produces including the removal renders, against the oracle on cuda."""
    rays = 1 << 15
    model, ocfg = pu.bench_like_model(tn, "separate", 19, "trained")
    model = model.to(DEV).eval()
    sd = {k: v.detach() for k, v in pu.oracle_state(model, DEV).items()}
    batch = _batch(rays, 43)
    with torch.no_grad():
        out = model(tn.RayBundle(origins=batch["origins"].clone(), directions=batch["directions"].clone(),
                                 pixel_area=batch["pixel_area"], camera_indices=batch["camera_indices"]))
        ref = oracle.thermal_nerfacto_forward(sd, ocfg, batch["origins"], batch["directions"], batch["camera_indices"],
                                              training=False)
    rep = {}
    for k in ("rgb", "rgb_thermal", "accumulation", "accumulation_thermal", "expected_depth", "expected_depth_thermal",
              "removal", "removal_thermal"):
        r = ref[k]
        d = (out[k] - r).abs()
        if "depth" in k:
            d = d / r.abs().clamp(min=1.0)
        rep[k] = d.max().item()
    for sfx in ("", "_thermal"):
        rep[f"depth{sfx}"] = pu.median_depth_check(out[f"depth{sfx}"], ref[f"_weights{sfx}"][-1], ref[f"_steps{sfx}"][-1],
                                                   strict=False)
    pu.report("eval_separate_32768_T19_trained", rep)
    # the removal renders threshold |1 - d2/d| < 0.05 per sample (thermal_nerfacto.py:463-485): a discontinuous mask,
    # so a handful of rays may legitimately flip one sample; everything else meets 1e-3 outright
    for k, v in rep.items():
        if isinstance(v, dict):
            assert v["unexcused"] == 0 and v["misplaced"] == 0, (k, v)
        elif not k.startswith("removal"):
            assert v <= pu.OUT_TOL, (k, v)
    for k in ("removal", "removal_thermal"):
        frac = ((out[k] - ref[k]).abs().max(dim=-1).values > pu.OUT_TOL).float().mean().item()
        assert frac <= 2e-3, (k, frac)
