"""Optimiser step on CPU: the oracle and the host-side schedule against the fixture produced by
the reference's own AdamOptimizerConfig / ExponentialDecayScheduler objects (tests/golden/make_golden_optim.py)."""
import numpy as np
import pytest
import torch

import oracle.optim as ooptim
from nerfstudio_thermal_b200 import optim as poptim
from nerfstudio_thermal_b200.parallel import FlatGradBuffer

GROUPS = {  # must match tests/golden/make_golden_optim.py
    "fields": (4, 1e-2, 1e-4, 200000, 0, "cosine"),
    "camera_opt": (1, 1e-3, 1e-4, 5000, 0, "cosine"),
    "warm": (2, 5e-3, 5e-5, 40, 6, "cosine"),
    "warm_linear": (1, 2e-3, None, 30, 4, "linear"),
}


def group_config(name):
    _, lr, lr_final, max_steps, warmup, ramp = GROUPS[name]
    return poptim.AdamGroupConfig(lr=lr, eps=1e-15, lr_final=lr_final, max_steps=max_steps, warmup_steps=warmup,
                                  ramp=ramp)


@pytest.mark.parametrize("name", list(GROUPS))
def test_schedule_matches_reference(golden, name):
    g = {k: v.numpy() for k, v in golden("optim.npz").items()}
    _, lr, lr_final, max_steps, warmup, ramp = GROUPS[name]
    probe = g["lr_probe_steps"]
    want = g[f"{name}_lr_probe"]
    got_oracle = np.array([ooptim.exponential_decay_lr(int(k), lr, lr_final, max_steps, warmup, 1e-8, ramp) for k in probe])
    got_host = np.array([poptim.scheduled_lr(group_config(name), int(k)) for k in probe])
    np.testing.assert_allclose(got_oracle, want, rtol=1e-14, atol=0)
    np.testing.assert_allclose(got_host, want, rtol=1e-14, atol=0)
    # the rate each Adam step actually used (step t uses lr_lambda(t-1))
    used = np.array([ooptim.exponential_decay_lr(t, lr, lr_final, max_steps, warmup, 1e-8, ramp)
                     for t in range(int(g["steps"]))])
    np.testing.assert_allclose(used, g[f"{name}_lrs"], rtol=1e-14, atol=0)


@pytest.mark.parametrize("name", list(GROUPS))
def test_oracle_adam_matches_reference(golden, name):
    g = {k: v.numpy() for k, v in golden("optim.npz").items()}
    n_params, lr, lr_final, max_steps, warmup, ramp = GROUPS[name]
    for i in range(n_params):
        p = torch.from_numpy(g[f"{name}_p{i}_init"]).clone()
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        for t in range(int(g["steps"])):
            lr_t = ooptim.exponential_decay_lr(t, lr, lr_final, max_steps, warmup, 1e-8, ramp)
            p, m, v = ooptim.adam_step(p, torch.from_numpy(g[f"{name}_p{i}_grad{t}"]), m, v, t + 1, lr_t)
        torch.testing.assert_close(p, torch.from_numpy(g[f"{name}_p{i}_final"]), rtol=2e-6, atol=1e-8)
        torch.testing.assert_close(m, torch.from_numpy(g[f"{name}_p{i}_exp_avg"]), rtol=2e-6, atol=1e-12)
        torch.testing.assert_close(v, torch.from_numpy(g[f"{name}_p{i}_exp_avg_sq"]), rtol=2e-6, atol=1e-14)


def test_method_config_table():
    """configs/method_configs.py:274-301."""
    cfg = poptim.thermal_nerfacto_optimizers()
    assert set(cfg) == {"proposal_networks", "fields", "proposal_networks_thermal", "fields_thermal", "camera_opt",
                        "camera_opt_thermal", "shared_camera_opt"}
    assert cfg["fields"].lr == 1e-2 and cfg["fields"].eps == 1e-15 and cfg["fields"].max_steps == 200000
    assert cfg["camera_opt"].lr == 1e-3 and cfg["camera_opt"].lr_final == 1e-4 and cfg["camera_opt"].max_steps == 5000


def test_flat_buffer_groups_and_flat_params():
    torch.manual_seed(0)
    a = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(1))]
    b = [torch.nn.Parameter(torch.randn(7)), a[0]]  # a[0] listed twice: stays in its first group
    buf = FlatGradBuffer.from_param_groups({"a": a, "b": b})
    assert buf.group_ranges == {"a": (0, 20), "b": (20, 28)}
    assert [off for _, off in buf.group_params("a")] == [0, 16] and [off for _, off in buf.group_params("b")] == [20]
    before = [p.detach().clone() for p in buf.params]
    flat = buf.flatten_params()
    assert flat.numel() == buf.flat.numel()
    for p, off, ref in zip(buf.params, buf.offsets, before):
        assert torch.equal(p.detach(), ref)
        assert p.data.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()
        assert torch.equal(flat[off:off + p.numel()].view_as(p), ref)
    assert buf.flatten_params() is flat
    assert buf.check_views()


def test_fused_adam_refuses_cpu():
    p = [torch.nn.Parameter(torch.randn(4))]
    buf = FlatGradBuffer.from_param_groups({"fields": p})
    with pytest.raises(RuntimeError, match="CUDA only"):
        poptim.FusedAdam(buf, poptim.thermal_nerfacto_optimizers())
    with pytest.raises(RuntimeError, match="not found"):
        poptim.FusedAdam(FlatGradBuffer.from_param_groups({"nope": p}), poptim.thermal_nerfacto_optimizers())
