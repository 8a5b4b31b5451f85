"""Golden fixture for the optimiser step: tests/golden/optim.npz.

Runs the UNMODIFIED reference classes -- `AdamOptimizerConfig.setup` (engine/optimizers.py:31-62) and
`ExponentialDecayScheduler.get_scheduler` (engine/schedulers.py:109-142) with the thermal-nerfacto settings of
configs/method_configs.py:274-301 -- on two small parameter groups for a few steps.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_optim.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_shim  # noqa: E402

_ref_shim.install()

from nerfstudio.engine.optimizers import AdamOptimizerConfig  # noqa: E402
from nerfstudio.engine.schedulers import ExponentialDecaySchedulerConfig  # noqa: E402

GROUPS = {  # name: (shapes, lr, lr_final, max_steps, warmup_steps, ramp)
    "fields": ([(37, 2), (16, 5), (16,), (1,)], 1e-2, 1e-4, 200000, 0, "cosine"),
    "camera_opt": ([(8, 6)], 1e-3, 1e-4, 5000, 0, "cosine"),
    "warm": ([(11,), (3, 3)], 5e-3, 5e-5, 40, 6, "cosine"),
    "warm_linear": ([(7,)], 2e-3, None, 30, 4, "linear"),
}
STEPS = 12
LR_PROBE = [0, 1, 2, 3, 5, 6, 7, 29, 30, 31, 39, 40, 41, 100, 4999, 5000, 5001, 100000, 200000, 250000]


def main():
    out = {}
    gen = torch.Generator().manual_seed(7)
    for name, (shapes, lr, lr_final, max_steps, warmup, ramp) in GROUPS.items():
        params = [torch.nn.Parameter(torch.randn(*s, generator=gen) * 0.1) for s in shapes]
        opt = AdamOptimizerConfig(lr=lr, eps=1e-15).setup(params=params)
        sched_cfg = ExponentialDecaySchedulerConfig(lr_final=lr_final, max_steps=max_steps, warmup_steps=warmup, ramp=ramp)
        sched = sched_cfg.setup().get_scheduler(optimizer=opt, lr_init=lr)
        out[f"{name}_lr_probe"] = np.array([lr * sched.lr_lambdas[0](k) for k in LR_PROBE], dtype=np.float64)
        for i, p in enumerate(params):
            out[f"{name}_p{i}_init"] = p.detach().numpy().copy()
        lrs = []
        for t in range(STEPS):
            for i, p in enumerate(params):
                g = torch.randn(p.shape, generator=gen) * (0.0 if (t == 3 and i == 0) else 1e-2)  # one all-zero grad
                p.grad = g
                out[f"{name}_p{i}_grad{t}"] = g.numpy().copy()
            lrs.append(opt.param_groups[0]["lr"])
            opt.step()
            sched.step()
        out[f"{name}_lrs"] = np.array(lrs, dtype=np.float64)
        for i, p in enumerate(params):
            out[f"{name}_p{i}_final"] = p.detach().numpy().copy()
            st = opt.state[p]
            out[f"{name}_p{i}_exp_avg"] = st["exp_avg"].numpy().copy()
            out[f"{name}_p{i}_exp_avg_sq"] = st["exp_avg_sq"].numpy().copy()
    out["lr_probe_steps"] = np.array(LR_PROBE)
    out["steps"] = np.array(STEPS)
    path = os.path.join(HERE, "optim.npz")
    np.savez_compressed(path, **out)
    print(f"wrote optim.npz: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
