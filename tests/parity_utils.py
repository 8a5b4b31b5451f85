"""Helpers of the GPU parity tests: build the bench-size model, run the oracle on the SAME device, compare.

The oracle (oracle/, pinned bit-for-bit against the unmodified reference on the CPU, tests/test_oracle_golden.py)
is device-agnostic torch code: run on `cuda` it is the north star's literal target, the reference's
`implementation="torch"` ops on the B200.  Tolerances (BASELINE.json north_star): rendered rgb / thermal / depth /
accumulation <= 1e-3 max abs, gradients <= 1e-3 relative (L2 per parameter tensor), sample counts exact.
"""
import json
import os
from typing import Dict, List, Optional

import torch

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT_DIR = os.environ.get("TN_PARITY_REPORT", os.path.join(ROOT, "gpurun_out"))

OUT_TOL = 1e-3   # max abs, rendered outputs
GRAD_TOL = 1e-3  # relative L2, per parameter tensor
LOSS_TOL = 1e-3  # relative, per loss term
MEDIAN_EPS = 1e-3  # |cumsum(w) - 0.5| allowed at a bin the median-depth search flipped across (= the accumulation bar)


def report(name: str, payload: dict) -> None:
    """Measured parity numbers, kept for DESIGN.md (gpurun_out/ is merged back from the GPU box)."""
    try:
        os.makedirs(REPORT_DIR, exist_ok=True)
        with open(os.path.join(REPORT_DIR, f"parity_{name}.json"), "w") as f:
            json.dump(payload, f, indent=1, sort_keys=True)
    except OSError:
        pass


def rel_l2(got: torch.Tensor, want: torch.Tensor) -> float:
    want = want.double()
    return ((got.double().to(want.device) - want).norm() / (want.norm() + 1e-300)).item()


def median_depth_check(depth_got: torch.Tensor, w_ref: torch.Tensor, steps_ref: torch.Tensor,
                       eps: float = MEDIAN_EPS, strict: bool = True) -> Dict[str, float]:
    """DepthRenderer(method="median") (renderers.py:547-557) picks the first sample whose cumulative weight reaches
    0.5: a discontinuous function of the weights.  Instead of a mismatch budget: a ray may land on another sample
    than the reference's ONLY if every cumulative weight between the two picks is within `eps` of 0.5 (the
    reference's own prefix sum at the bins the search flipped across).  Returns the observed statistics and raises
    on a ray that differs without that excuse.

    depth_got [R,1]; w_ref [R,S,1] reference weights; steps_ref [R,S,1] reference sample midpoints."""
    w = w_ref[..., 0].double()
    steps = steps_ref[..., 0]
    R, S = w.shape
    cum = torch.cumsum(w_ref[..., 0], dim=-1)  # float32, as the reference
    i_ref = torch.clamp(torch.searchsorted(cum, torch.full((R, 1), 0.5, device=cum.device), side="left"), 0, S - 1)[:, 0]
    got = depth_got.reshape(R).to(steps.device)
    # the sample our depth sits on: nearest reference midpoint (sample placement itself is compared elsewhere)
    j_got = (steps - got[:, None]).abs().argmin(dim=1)
    # (far samples: euclid = 1/(2-2s) turns a 4e-7 difference of a spacing-domain bin into ~1e-3 relative at
    # depth 1000, so "sits on that sample" is judged at 5e-3 relative)
    placed = (steps.gather(1, j_got[:, None])[:, 0] - got).abs() <= 5e-3 * got.abs().clamp(min=1.0)
    differs = j_got != i_ref
    lo, hi = torch.minimum(i_ref, j_got), torch.maximum(i_ref, j_got)
    ar = torch.arange(S, device=cum.device)[None, :]
    between = (ar >= lo[:, None]) & (ar < hi[:, None])
    dist = torch.where(between, (cum - 0.5).abs(), torch.zeros_like(cum)).max(dim=1).values
    bad = differs & (dist > eps)
    stats = {"rays": R, "differ_frac": differs.float().mean().item(), "max_boundary_dist": dist.max().item(),
             "unexcused": int(bad.sum().item()), "misplaced": int((~placed).sum().item())}
    if strict:
        assert stats["misplaced"] == 0, f"median depth not on a reference sample midpoint: {stats}"
        assert stats["unexcused"] == 0, f"median depth differs away from the 0.5 boundary: {stats}"
    return stats


def bench_like_model(tn, mode: str, log2_T: int, init: str, num_cams: int = 64, seed: int = 1234):
    """The model bench.py builds (build_model): default sizes, 64 cameras (32 RGB + 32 thermal), and either the
    reference's initialisers or the 'trained-like' tables of SURVEY.md 8(d)."""
    torch.manual_seed(seed)
    cfg = tn.ThermalNerfactoModelConfig(density_mode=mode, log2_hashmap_size=log2_T)
    half = num_cams // 2
    model = cfg.setup(num_train_data=num_cams, metadata={"is_thermal": [0] * half + [1] * (num_cams - half)})
    with torch.no_grad():
        for k, p in model.named_parameters():
            if "pose_adjustment" in k:
                p.normal_(0, 1e-3)  # zeros would sit on the kink of |t| in the regulariser
        if init == "trained":
            for k, p in model.named_parameters():
                if k.endswith("hash_table"):
                    p.uniform_(-0.5, 0.5)
            for f in [model.field] + ([model.field_thermal] if mode == "separate" else []):
                f.mlp_base.model[1].layers[-1].bias[0] += 2.0
    ocfg = oracle.OracleConfig(density_mode=mode, log2_hashmap_size=log2_T,
                               is_thermal_cameras=tuple([0] * half + [1] * (num_cams - half)))
    return model, ocfg


def oracle_state(model, device) -> Dict[str, torch.Tensor]:
    """The model's state_dict as oracle leaves on `device` (aliased proposal tables are one tensor)."""
    sd = {}
    for k, v in model.state_dict().items():
        v = v.detach().clone().to(device)
        if v.dtype == torch.float32 and v.numel() > 0 and not k.endswith("aabb") and v.dim() > 0:
            v.requires_grad_(True)
        sd[k] = v
    for k in list(sd):
        if k.endswith("encoding.hash_table"):
            sd[k] = sd[k.replace("encoding.hash_table", "mlp_base.0.hash_table")]
    return sd


def oracle_train_step(sd, ocfg, batch, jit: List[torch.Tensor], anneal: float = 1.0, updated: bool = True):
    for v in sd.values():
        if v.requires_grad:
            v.grad = None
    sep = ocfg.density_mode == "separate"
    out = oracle.thermal_nerfacto_forward(sd, ocfg, batch["origins"], batch["directions"], batch["camera_indices"],
                                          training=True, jitters=jit[:3], jitters_thermal=jit[3:] if sep else None,
                                          anneal=anneal, updated=updated)
    losses = oracle.thermal_nerfacto_losses(sd, ocfg, out, batch["image"], batch["is_thermal"], training=True)
    sum(losses.values()).backward()
    return out, losses


def param_grads_vs_oracle(model, sd) -> Dict[str, float]:
    """relative L2 error of every parameter gradient the oracle produced (keyed by the model's parameter name)."""
    res = {}
    for k, p in model.named_parameters():
        kk = k.replace("encoding.hash_table", "mlp_base.0.hash_table")
        want = sd[kk].grad if kk in sd else None
        if want is None or float(want.abs().max()) < 1e-12:
            continue
        assert p.grad is not None, k
        res[k] = rel_l2(p.grad, want)
    return res


def compare_outputs(out, ref, mode: str, strict: bool = True) -> Dict[str, float]:
    """max abs error of the rendered outputs + the median-depth boundary check (train-mode output dicts)."""
    errs = {}
    keys = ["rgb", "accumulation", "expected_depth"]
    if mode != "rgb_only":
        keys.append("rgb_thermal")
    if mode == "separate":
        keys += ["accumulation_thermal", "expected_depth_thermal"]
    for k in keys:
        r = ref[k].detach().to(out[k].device)
        d = (out[k].detach() - r).abs()
        if "depth" in k:  # depths reach far=1000 (fp32 ulp 6e-5 there): 1e-3 absolute up to 1, relative beyond
            d = d / r.abs().clamp(min=1.0)
        errs[k] = d.max().item()
    for sfx in (("", "_thermal") if mode == "separate" else ("",)):
        w, st = ref[f"_weights{sfx}"], ref[f"_steps{sfx}"]
        errs[f"depth{sfx}"] = median_depth_check(out[f"depth{sfx}"].detach(), w[-1], st[-1], strict=strict)
        for i in range(len(w) - 1):
            errs[f"prop_depth_{i}{sfx}"] = median_depth_check(out[f"prop_depth_{i}{sfx}"].detach(), w[i], st[i], strict=strict)
    return errs
