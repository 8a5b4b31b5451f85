"""Optimiser step of the hot path: one fused, graph-capturable Adam launch over the flat buffers.

Mirrors the reference's `Optimizers` wrapper (engine/optimizers.py:67-210) for what this model uses: one
`torch.optim.Adam` (eps 1e-15, no weight decay, no clipping) plus one `ExponentialDecayScheduler`
(engine/schedulers.py:109-142) per parameter group, groups and hyper-parameters as in
configs/method_configs.py:274-301.  Instead of G optimizers x (foreach kernels per state tensor) the parameters,
gradients and both moments live in flat fp32 buffers laid out group by group (`parallel.FlatGradBuffer`), and
`csrc/tn_optim.cu` updates all groups in one streaming kernel that also evaluates each group's learning rate from
a device-side step counter -- so the whole train iteration, optimiser included, replays as one CUDA graph.
"""
from ctypes import c_int64
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from ._lib import call, float_array, ptr, stream
from .parallel import FlatGradBuffer


@dataclass
class AdamGroupConfig:
    """AdamOptimizerConfig + ExponentialDecaySchedulerConfig of one group (optimizers.py:31-62, schedulers.py:92-107)."""

    lr: float = 1e-2
    eps: float = 1e-15
    weight_decay: float = 0.0
    max_norm: Optional[float] = None
    lr_final: Optional[float] = None
    lr_pre_warmup: float = 1e-8
    warmup_steps: int = 0
    max_steps: int = 0  # 0: no scheduler (constant lr)
    ramp: str = "cosine"


def thermal_nerfacto_optimizers() -> Dict[str, AdamGroupConfig]:
    """The `optimizers=` table of the thermal-nerfacto method (configs/method_configs.py:274-301)."""
    net = dict(lr=1e-2, eps=1e-15, lr_final=1e-4, max_steps=200000)
    cam = dict(lr=1e-3, eps=1e-15, lr_final=1e-4, max_steps=5000)
    return {
        "proposal_networks": AdamGroupConfig(**net), "fields": AdamGroupConfig(**net),
        "proposal_networks_thermal": AdamGroupConfig(**net), "fields_thermal": AdamGroupConfig(**net),
        "camera_opt": AdamGroupConfig(**cam), "camera_opt_thermal": AdamGroupConfig(**cam),
        "shared_camera_opt": AdamGroupConfig(**cam),
    }


def scheduled_lr(cfg: AdamGroupConfig, k: int) -> float:
    """Learning rate after k scheduler steps (host copy of the kernel's closed form, for logging and state dicts):
    lr_init * lr_lambda(k), schedulers.py:124-139."""
    if cfg.max_steps <= 0:
        return cfg.lr
    lr_final = cfg.lr if cfg.lr_final is None else cfg.lr_final
    if k < cfg.warmup_steps:
        if cfg.ramp == "cosine":
            lr = cfg.lr_pre_warmup + (cfg.lr - cfg.lr_pre_warmup) * np.sin(
                0.5 * np.pi * np.clip(k / cfg.warmup_steps, 0, 1))
        else:
            lr = cfg.lr_pre_warmup + (cfg.lr - cfg.lr_pre_warmup) * k / cfg.warmup_steps
    else:
        t = np.clip((k - cfg.warmup_steps) / (cfg.max_steps - cfg.warmup_steps), 0, 1)
        lr = np.exp(np.log(cfg.lr) * (1 - t) + np.log(lr_final) * t)
    return float(cfg.lr * (lr / cfg.lr))


class FusedAdam:
    """All parameter groups of a `FlatGradBuffer` stepped by one kernel launch.

    step() = for every group: `optimizer.step(); scheduler.step()` (trainer.py:476-492 order: the k-th Adam step
    uses lr_lambda(k-1)).  Gradients are read from `grads.flat`; parameters are moved into one flat buffer
    (`param.data` become views of it) so that element i of every buffer belongs to the same scalar.
    """

    def __init__(self, grads: FlatGradBuffer, config: Dict[str, AdamGroupConfig],
                 betas: Tuple[float, float] = (0.9, 0.999)):
        if not grads.group_ranges:
            raise ValueError("FusedAdam needs a FlatGradBuffer built with from_param_groups (named groups)")
        missing = [n for n in grads.group_ranges if n not in config]
        if missing:  # optimizers.py:95-98
            raise RuntimeError(f"Optimizer config for {missing} not found; provided configs were: {list(config)}")
        if grads.flat.device.type != "cuda":
            raise RuntimeError("FusedAdam runs on CUDA only (no CPU fallback)")
        self.grads, self.betas = grads, betas
        self.names = list(grads.group_ranges)
        self.config = {n: config[n] for n in self.names}
        for n, c in self.config.items():
            if c.max_norm is not None:
                raise NotImplementedError(f"group {n}: gradient clipping (max_norm) is not part of this path")
        self.params = grads.flatten_params()
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        g = len(self.names)
        # device-side counters read by the kernel (capturable): [0] iterations = scheduler steps, [1+k] Adam steps of
        # group k (torch's state['step']; lags the iteration count for groups that sat out steps without a gradient)
        self.step_dev = torch.zeros(1 + g, dtype=torch.int32, device=self.params.device)
        self._all_active = (1 << g) - 1
        self._active = self._all_active
        self._begin = (c_int64 * g)(*[grads.group_ranges[n][0] for n in self.names])
        self._end = (c_int64 * g)(*[grads.group_ranges[n][1] for n in self.names])
        hyper: List[float] = []
        for n in self.names:
            c = self.config[n]
            hyper += [c.lr, -1.0 if c.lr_final is None else c.lr_final, c.lr_pre_warmup, c.eps, c.weight_decay,
                      c.warmup_steps, c.max_steps, 1.0 if c.ramp == "cosine" else 0.0]
        self._hyper = float_array(hyper)

    @property
    def step_count(self) -> int:
        """Iterations taken (= scheduler steps)."""
        return int(self.step_dev[0].item())

    def group_step_counts(self) -> Dict[str, int]:
        """Adam steps taken per group (torch's state['step'])."""
        v = self.step_dev.tolist()
        return {n: int(v[1 + i]) for i, n in enumerate(self.names)}

    def active_mask(self, inactive: Optional[List[str]] = None) -> int:
        """Bit mask of the groups that take part in a step; `inactive` names the groups without a gradient this
        iteration (the reference skips them: torch.optim.Adam ignores parameters whose grad is None and
        Optimizers.optimizer_scaler_step_all skips optimisers with no gradient at all, optimizers.py:150-170)."""
        mask = self._all_active
        for n in inactive or ():
            if n in self.names:
                mask &= ~(1 << self.names.index(n))
        return mask

    def step(self, zero_grads: bool = False, inv_scale: Optional[Tensor] = None,
             found_inf: Optional[Tensor] = None, inactive: Optional[List[str]] = None) -> None:
        """One optimiser + scheduler step of every group.  inv_scale / found_inf: the GradScaler's device scalars
        (optimizers.py:150-163); `zero_grads` clears the gradient buffer in the same pass (zero_grad_all);
        `inactive`: groups without a gradient this iteration (left untouched, their step count does not advance;
        their scheduler still steps, as in Trainer.train_iteration)."""
        self.tick(inactive)
        self.step_range(0, self.params.numel(), zero_grads, inv_scale, found_inf)

    def tick(self, inactive: Optional[List[str]] = None) -> None:
        """Advance the device-side counters (once per iteration, before any step_range of that iteration) and fix
        which groups the step_range calls of this iteration update."""
        self._active = self.active_mask(inactive)
        call("tn_step_counters_tick", ptr(self.step_dev), len(self.names), self._active, stream())

    def step_range(self, begin: int, end: int, zero_grads: bool = False, inv_scale: Optional[Tensor] = None,
                   found_inf: Optional[Tensor] = None) -> None:
        """The update of elements [begin, end) of the flat buffers only (begin a multiple of 4): lets a runner step
        the parameters whose gradients are already final while the rest of the backward is still running."""
        if begin % 4 != 0 or not 0 <= begin <= end <= self.params.numel():
            raise ValueError(f"bad range [{begin}, {end})")
        n = end - begin
        if n == 0:
            return
        g = len(self.names)
        clamp = lambda v: min(max(v - begin, 0), n)  # noqa: E731
        b = (c_int64 * g)(*[clamp(self.grads.group_ranges[k][0]) for k in self.names])
        e = (c_int64 * g)(*[clamp(self.grads.group_ranges[k][1]) for k in self.names])
        off = begin * 4
        call("tn_adam_step", ptr(self.params) + off, ptr(self.grads.flat) + off, ptr(self.exp_avg) + off,
             ptr(self.exp_avg_sq) + off, n, b, e, self._hyper, g, float(self.betas[0]), float(self.betas[1]),
             ptr(self.step_dev), 0, 1, self._active, ptr(inv_scale), ptr(found_inf), int(zero_grads), stream())

    def unscale_and_check(self, inv_scale: Optional[Tensor], found_inf: Tensor) -> None:
        """GradScaler.unscale_'s inf/nan check over the whole gradient buffer (the scaling itself is applied inside
        step())."""
        call("tn_grad_unscale_check", ptr(self.grads.flat), self.grads.flat.numel(), ptr(inv_scale), ptr(found_inf),
             stream())

    def get_last_lr(self) -> Dict[str, float]:
        """Per-group learning rate the NEXT step will use (scheduler.get_last_lr(), optimizers.py:190-192)."""
        k = self.step_count
        return {n: scheduled_lr(self.config[n], k) for n in self.names}

    # ---- checkpoint compatibility with the reference's per-group torch.optim.Adam state (trainer.py:389-453)
    def state_dict(self) -> Dict[str, dict]:
        k = self.step_count
        steps = self.group_step_counts()
        out = {}
        for n in self.names:
            state, idx = {}, []
            for i, (p, off) in enumerate(self.grads.group_params(n)):
                sl = slice(off, off + p.numel())
                state[i] = {"step": torch.tensor(float(steps[n])), "exp_avg": self.exp_avg[sl].view_as(p).clone(),
                            "exp_avg_sq": self.exp_avg_sq[sl].view_as(p).clone()}
                idx.append(i)
            c = self.config[n]
            out[n] = {"state": state if steps[n] > 0 else {},
                      "param_groups": [{"lr": scheduled_lr(c, k), "betas": self.betas, "eps": c.eps,
                                        "weight_decay": c.weight_decay, "amsgrad": False, "maximize": False,
                                        "initial_lr": c.lr, "params": idx}]}
        return out

    def load_state_dict(self, loaded: Dict[str, dict], schedulers: Optional[Dict[str, dict]] = None) -> None:
        """Optimizers.load_optimizers / load_schedulers (optimizers.py:194-210): adopt the moments, every group's own
        step count (real reference checkpoints disagree across groups: the proposal networks sit out the iterations
        on which they are not updated) and the schedulers' iteration count (`last_epoch`; without `schedulers` the
        largest group step stands in for it)."""
        counts = self.step_dev.tolist()
        for n, sd in loaded.items():
            if n not in self.config:
                raise KeyError(f"unknown parameter group {n}")
            gi = self.names.index(n)
            group_steps = set()
            for i, (p, off) in enumerate(self.grads.group_params(n)):
                st = sd["state"].get(i)
                if st is None:
                    continue
                sl = slice(off, off + p.numel())
                self.exp_avg[sl].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[sl].copy_(st["exp_avg_sq"].reshape(-1))
                group_steps.add(int(float(st["step"])))
            if len(group_steps) > 1:
                # one counter per group: parameters of a group that disagree cannot be represented
                raise ValueError(f"group {n}: parameters disagree on the step count: {sorted(group_steps)}")
            counts[1 + gi] = group_steps.pop() if group_steps else 0
        epochs = [int(v["last_epoch"]) for k, v in (schedulers or {}).items() if k in self.config and "last_epoch" in v]
        counts[0] = max(epochs) if epochs else max(counts[1:], default=0)
        self.step_dev.copy_(torch.tensor(counts, dtype=torch.int32))
