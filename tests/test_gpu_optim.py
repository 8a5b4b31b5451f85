"""Fused Adam kernel (csrc/tn_optim.cu) through the C ABI: against the reference fixture, the oracle and torch."""
import numpy as np
import pytest
import torch

import oracle.optim as ooptim
from nerfstudio_thermal_b200 import optim as poptim
from nerfstudio_thermal_b200.parallel import FlatGradBuffer

from test_optim_cpu import GROUPS, group_config

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_fused_adam_matches_reference_fixture(golden):
    """All four groups of the fixture stepped by ONE launch per step; final p / exp_avg / exp_avg_sq are those of
    the reference's torch.optim.Adam + ExponentialDecayScheduler objects."""
    g = golden("optim.npz")
    groups = {n: [torch.nn.Parameter(g[f"{n}_p{i}_init"].clone().to(DEV)) for i in range(GROUPS[n][0])] for n in GROUPS}
    buf = FlatGradBuffer.from_param_groups(groups)
    opt = poptim.FusedAdam(buf, {n: group_config(n) for n in GROUPS})
    steps = int(g["steps"])
    for t in range(steps):
        for n, ps in groups.items():
            for i, p in enumerate(ps):
                p.grad.copy_(g[f"{n}_p{i}_grad{t}"])
        lrs = opt.get_last_lr()
        for n in GROUPS:
            assert lrs[n] == pytest.approx(float(g[f"{n}_lrs"][t]), rel=1e-14)
        opt.step()
    assert opt.step_count == steps
    for n, ps in groups.items():
        for i, (p, off) in enumerate(buf.group_params(n)):
            sl = slice(off, off + p.numel())
            torch.testing.assert_close(p.detach().cpu(), g[f"{n}_p{i}_final"], rtol=1e-5, atol=1e-7)
            torch.testing.assert_close(opt.exp_avg[sl].view_as(p).cpu(), g[f"{n}_p{i}_exp_avg"], rtol=1e-5, atol=1e-10)
            torch.testing.assert_close(opt.exp_avg_sq[sl].view_as(p).cpu(), g[f"{n}_p{i}_exp_avg_sq"], rtol=1e-5, atol=1e-12)


def _random_groups(seed, shapes):
    gen = torch.Generator().manual_seed(seed)
    return {n: [torch.nn.Parameter((torch.randn(*s, generator=gen) * 0.2).to(DEV)) for s in ss] for n, ss in shapes.items()}


def test_fused_adam_vs_oracle_ragged_groups_weight_decay():
    """Group boundaries inside a float4, a million-element table, weight decay, zero_grads."""
    shapes = {"a": [(1 << 20, 2), (64, 32), (64,), (1, 64), (1,)], "b": [(3,), (5, 7), (2,)], "c": [(9,)]}
    groups = _random_groups(1, shapes)
    cfg = {"a": poptim.AdamGroupConfig(lr=1e-2, lr_final=1e-4, max_steps=100),
           "b": poptim.AdamGroupConfig(lr=3e-3, weight_decay=0.1, eps=1e-8),
           "c": poptim.AdamGroupConfig(lr=1e-3, lr_final=1e-5, max_steps=50, warmup_steps=3)}
    buf = FlatGradBuffer.from_param_groups(groups)
    ref = {id(p): [p.detach().cpu().clone(), torch.zeros(p.shape), torch.zeros(p.shape)] for ps in groups.values() for p in ps}
    opt = poptim.FusedAdam(buf, cfg)
    gen = torch.Generator().manual_seed(2)
    for t in range(5):
        for n, ps in groups.items():
            c = cfg[n]
            lr = ooptim.exponential_decay_lr(t, c.lr, c.lr_final, c.max_steps, c.warmup_steps, c.lr_pre_warmup, c.ramp) \
                if c.max_steps else c.lr
            for p in ps:
                gr = torch.randn(p.shape, generator=gen) * 10.0 ** float(torch.randint(-6, 1, (1,), generator=gen))
                p.grad.copy_(gr)
                st = ref[id(p)]
                st[0], st[1], st[2] = ooptim.adam_step(st[0], gr, st[1], st[2], t + 1, lr, c.eps, c.weight_decay)
        opt.step(zero_grads=(t == 4))
    assert torch.count_nonzero(buf.flat) == 0  # cleared by the last step
    for n, ps in groups.items():
        for p, off in buf.group_params(n):
            torch.testing.assert_close(p.detach().cpu(), ref[id(p)][0], rtol=1e-5, atol=1e-7)
            torch.testing.assert_close(opt.exp_avg_sq[off:off + p.numel()].view_as(p).cpu(), ref[id(p)][2], rtol=1e-5, atol=1e-14)


def test_fused_adam_matches_torch_adam_and_state_dict_roundtrip():
    shapes = {"fields": [(1000, 2), (16, 10), (16,)], "camera_opt": [(8, 6)]}
    groups = _random_groups(3, shapes)
    twins = {n: [torch.nn.Parameter(p.detach().clone()) for p in ps] for n, ps in groups.items()}
    cfg = poptim.thermal_nerfacto_optimizers()
    buf = FlatGradBuffer.from_param_groups(groups)
    opt = poptim.FusedAdam(buf, {n: cfg[n] for n in shapes})
    topt = {n: torch.optim.Adam(ps, lr=cfg[n].lr, eps=cfg[n].eps) for n, ps in twins.items()}
    gen = torch.Generator().manual_seed(4)

    def both_step():
        for n in shapes:
            for p, q in zip(groups[n], twins[n]):
                gr = (torch.randn(p.shape, generator=gen) * 1e-3).to(DEV)
                p.grad.copy_(gr)
                q.grad = gr.clone()
            for gpar in topt[n].param_groups:
                gpar["lr"] = opt.get_last_lr()[n]
            topt[n].step()
        opt.step()

    for _ in range(4):
        both_step()
    for n in shapes:
        for p, q in zip(groups[n], twins[n]):
            torch.testing.assert_close(p.detach(), q.detach(), rtol=1e-5, atol=1e-7)
    # our state dict loads into the reference's optimizer objects and back (trainer.py:389-453 checkpoints)
    sd = opt.state_dict()
    for n in shapes:
        fresh = torch.optim.Adam(twins[n], lr=cfg[n].lr, eps=cfg[n].eps)
        fresh.load_state_dict(sd[n])
        for q, p in zip(twins[n], groups[n]):
            torch.testing.assert_close(fresh.state[q]["exp_avg"], topt[n].state[q]["exp_avg"], rtol=1e-5, atol=1e-10)
            assert float(fresh.state[q]["step"]) == 4.0
    opt2_groups = {n: [torch.nn.Parameter(p.detach().clone()) for p in ps] for n, ps in groups.items()}
    buf2 = FlatGradBuffer.from_param_groups(opt2_groups)
    opt2 = poptim.FusedAdam(buf2, {n: cfg[n] for n in shapes})
    opt2.load_state_dict({n: topt[n].state_dict() for n in shapes})
    assert opt2.step_count == 4
    torch.testing.assert_close(opt2.exp_avg, opt.exp_avg, rtol=1e-5, atol=1e-10)
    torch.testing.assert_close(opt2.exp_avg_sq, opt.exp_avg_sq, rtol=1e-5, atol=1e-14)


def test_grad_scaler_path_skips_on_inf():
    groups = _random_groups(5, {"fields": [(257,), (3, 3)]})
    buf = FlatGradBuffer.from_param_groups(groups)
    opt = poptim.FusedAdam(buf, poptim.thermal_nerfacto_optimizers() | {})
    before = opt.params.clone()
    inv_scale = torch.tensor([1.0 / 1024.0], device=DEV)
    found = torch.zeros(1, device=DEV)
    buf.flat.normal_()
    buf.flat[100] = float("inf")
    opt.unscale_and_check(inv_scale, found)
    assert found.item() == 1.0
    opt.step(inv_scale=inv_scale, found_inf=found)
    assert torch.equal(opt.params, before) and torch.count_nonzero(opt.exp_avg) == 0
    buf.flat.normal_()
    opt.unscale_and_check(inv_scale, found)
    assert found.item() == 0.0
    g = buf.flat.clone()
    opt.step(inv_scale=inv_scale, found_inf=found)
    # first Adam step with scaled gradients: m = (1-b1) * g/1024
    torch.testing.assert_close(opt.exp_avg, 0.1 * g / 1024.0, rtol=1e-5, atol=1e-12)
    assert not torch.equal(opt.params, before)


def test_adam_step_argument_errors():
    from ctypes import c_int64

    from nerfstudio_thermal_b200 import TnKernelError
    from nerfstudio_thermal_b200._lib import call, float_array, ptr, stream
    x = torch.zeros(16, device=DEV)
    hyper = float_array([1e-2, -1, 1e-8, 1e-15, 0, 0, 0, 1])
    with pytest.raises(TnKernelError, match="range"):
        call("tn_adam_step", ptr(x), ptr(x), ptr(x), ptr(x), 16, (c_int64 * 1)(0), (c_int64 * 1)(17), hyper, 1, 0.9,
             0.999, None, 1, 0, 1, None, None, 0, stream())
    with pytest.raises(TnKernelError, match="step"):
        call("tn_adam_step", ptr(x), ptr(x), ptr(x), ptr(x), 16, (c_int64 * 1)(0), (c_int64 * 1)(16), hyper, 1, 0.9,
             0.999, None, 0, 0, 1, None, None, 0, stream())


def _small_model_and_batch(golden):
    import nerfstudio_thermal_b200 as tn
    from test_gpu_model import build
    g, model = build(golden, "separate")
    R = g["origins"].shape[0]
    batch = {"origins": g["origins"], "directions": g["directions"], "pixel_area": torch.full((R, 1), 1e-6),
             "camera_indices": g["camera_indices"], "image": g["image"], "is_thermal": g["is_thermal"]}
    return tn, model, {k: v.to(DEV) for k, v in batch.items()}


@pytest.mark.parametrize("use_graph", [True, False])
def test_train_step_with_fused_optimizer(golden, use_graph):
    """forward + loss + backward + Adam as one replayed CUDA graph: parameters move, the loss on a fixed batch goes
    down, warm-up/capture take no optimiser steps, gradients are left cleared by the in-graph Adam pass."""
    from nerfstudio_thermal_b200 import engine
    tn, model, batch = _small_model_and_batch(golden)
    p0 = {k: v.detach().clone() for k, v in model.named_parameters()}
    runner = engine.GraphedTrainStep(model, batch, use_graph=use_graph, optimizer=poptim.thermal_nerfacto_optimizers())
    assert runner.optimizer.step_count == 0
    for k, v in model.named_parameters():  # flattening and capture kept every value
        assert torch.equal(v.detach(), p0[k]), k
    losses = [float(runner.step(batch)) for _ in range(8)]
    assert runner.optimizer.step_count == 8
    assert losses[-1] < losses[0], losses
    moved = sum(int(not torch.equal(v.detach(), p0[k])) for k, v in model.named_parameters())
    assert moved >= 10
    if use_graph:
        assert torch.count_nonzero(runner.grads.flat) == 0
    if use_graph:  # a captured graph draws its jitter from the graph-registered Philox state: not comparable
        return
    # same model, same batch, torch.optim.Adam driven from param.grad: the first step moves parameters identically
    tn2, model2, _ = _small_model_and_batch(golden)
    runner2 = engine.GraphedTrainStep(model2, batch, use_graph=False)
    groups = model2.get_param_groups()
    cfg = poptim.thermal_nerfacto_optimizers()
    topt = [torch.optim.Adam(ps, lr=cfg[n].lr, eps=cfg[n].eps) for n, ps in groups.items()]
    tn3, model3, _ = _small_model_and_batch(golden)
    runner3 = engine.GraphedTrainStep(model3, batch, use_graph=use_graph, optimizer=cfg)
    torch.manual_seed(11); runner2.step(batch)
    for o in topt:
        o.step()
    torch.manual_seed(11); runner3.step(batch)
    p2, p3 = dict(model2.named_parameters()), dict(model3.named_parameters())
    for k in p2:
        # Adam's first step is lr * sign(g) wherever |g| >> eps: compare where the gradient is not at noise level
        g2 = p2[k].grad
        if g2 is None:
            continue
        sel = g2.abs() > 1e-7
        if sel.any():
            torch.testing.assert_close(p3[k].detach()[sel], p2[k].detach()[sel], rtol=1e-4, atol=2e-5)


def test_two_stream_branches_give_the_same_gradients_in_a_captured_step(golden):
    """The thermal branch on a second stream (model.branch_streams: parallel arms of the captured graph, forward and
    backward) must not change anything: same seed -> same jitter draws -> same losses and gradients as one stream."""
    from nerfstudio_thermal_b200 import engine
    res = []
    for two in (False, True):
        tn, model, batch = _small_model_and_batch(golden)
        model.branch_streams = two
        torch.manual_seed(77)
        runner = engine.GraphedTrainStep(model, batch, use_graph=True)
        for _ in range(3):
            total = runner.step(batch)
        res.append((float(total.detach()), runner.grads.flat.clone(), {k: float(v.detach()) for k, v in runner.losses.items()}))
    (t0, g0, l0), (t1, g1, l1) = res
    assert l0.keys() == l1.keys()
    for k in l0:
        assert l1[k] == pytest.approx(l0[k], rel=1e-5, abs=1e-9), k
    rel = ((g0 - g1).double().norm() / g0.double().norm()).item()
    assert rel <= 1e-5, rel


def test_inactive_groups_follow_torch_adam_with_grad_none():
    """A group without a gradient is skipped exactly like torch.optim.Adam skips parameters whose grad is None
    (engine/optimizers.py:150-170): its parameters and moments stay, its step count does not advance, its
    scheduler still steps; per-group step counts drive the bias corrections; state_dict / load_state_dict keep
    groups' differing step counts (a real reference checkpoint has them: ADVICE r1)."""
    shapes = {"a": [(33, 2), (7,)], "b": [(5, 3)], "c": [(9,)]}
    groups = _random_groups(5, shapes)
    cfg = {"a": poptim.AdamGroupConfig(lr=1e-2, lr_final=1e-4, max_steps=100),
           "b": poptim.AdamGroupConfig(lr=3e-3, lr_final=1e-4, max_steps=50),
           "c": poptim.AdamGroupConfig(lr=1e-3)}
    buf = FlatGradBuffer.from_param_groups(groups)
    opt = poptim.FusedAdam(buf, cfg)
    tgroups = {n: [torch.nn.Parameter(p.detach().clone()) for p in ps] for n, ps in groups.items()}
    topt = {n: torch.optim.Adam(ps, lr=cfg[n].lr, eps=cfg[n].eps) for n, ps in tgroups.items()}
    sched = {n: torch.optim.lr_scheduler.LambdaLR(
        topt[n], lambda k, c=cfg[n]: poptim.scheduled_lr(c, k) / c.lr) for n in topt}
    gen = torch.Generator().manual_seed(6)
    pattern = [[], ["b"], ["b"], [], ["b", "c"], []]  # groups without a gradient per iteration
    for inactive in pattern:
        for n, ps in groups.items():
            for p, tp in zip(ps, tgroups[n]):
                gr = torch.randn(p.shape, generator=gen).to(DEV)
                p.grad.copy_(gr)
                tp.grad = None if n in inactive else gr.clone()
        opt.step(inactive=inactive)
        for n in topt:
            if n not in inactive:
                topt[n].step()
            sched[n].step()
    assert opt.step_count == len(pattern)
    assert opt.group_step_counts() == {"a": 6, "b": 3, "c": 5}
    for n, ps in groups.items():
        for p, tp in zip(ps, tgroups[n]):
            torch.testing.assert_close(p.detach(), tp.detach(), rtol=2e-5, atol=1e-7)
    sd = opt.state_dict()
    assert float(sd["b"]["state"][0]["step"]) == 3.0 and float(sd["a"]["state"][0]["step"]) == 6.0
    for n in topt:  # the reference's own optimisers accept it, and the moments agree
        tsd = topt[n].state_dict()
        for i in tsd["state"]:
            torch.testing.assert_close(sd[n]["state"][i]["exp_avg"], tsd["state"][i]["exp_avg"], rtol=2e-5, atol=1e-9)
            assert float(tsd["state"][i]["step"]) == float(sd[n]["state"][i]["step"])
    # a fresh optimiser adopts differing step counts and continues identically
    groups2 = {n: [torch.nn.Parameter(p.detach().clone()) for p in ps] for n, ps in groups.items()}
    buf2 = FlatGradBuffer.from_param_groups(groups2)
    opt2 = poptim.FusedAdam(buf2, cfg)
    opt2.load_state_dict({n: topt[n].state_dict() for n in topt},
                         schedulers={n: sched[n].state_dict() for n in sched})
    assert opt2.step_count == len(pattern) and opt2.group_step_counts() == opt.group_step_counts()
    buf.flat.normal_(generator=None)
    buf2.flat.copy_(buf.flat)
    opt.step(); opt2.step()
    torch.testing.assert_close(opt2.params, opt.params, rtol=0, atol=0)
