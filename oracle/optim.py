"""Oracle: the optimiser step of the reference trainer.  TEST INFRASTRUCTURE ONLY.

The reference drives one `torch.optim.Adam` and one `LambdaLR` per parameter group
(engine/optimizers.py:67-95, 172-180; engine/schedulers.py:109-142; groups in configs/method_configs.py:274-301).
`torch.optim.Adam` is third-party code (torch 2.x, torch/optim/adam.py `_single_tensor_adam`, amsgrad/maximize/
capturable off); its published update rule is restated here.  Pinned by tests/golden/optim.npz, produced by the
reference's own `AdamOptimizerConfig.setup` / `ExponentialDecayScheduler.get_scheduler` objects
(tests/golden/make_golden_optim.py).
"""
import math
from typing import Optional, Tuple

import numpy as np
import torch


def exponential_decay_lr(k: int, lr_init: float, lr_final: Optional[float], max_steps: int, warmup_steps: int = 0,
                         lr_pre_warmup: float = 1e-8, ramp: str = "cosine") -> float:
    """Learning rate after k scheduler steps: lr_init * lr_lambda(k), engine/schedulers.py:124-139."""
    if lr_final is None:
        lr_final = lr_init
    if k < warmup_steps:
        if ramp == "cosine":
            lr = lr_pre_warmup + (lr_init - lr_pre_warmup) * np.sin(0.5 * np.pi * np.clip(k / warmup_steps, 0, 1))
        else:
            lr = lr_pre_warmup + (lr_init - lr_pre_warmup) * k / warmup_steps
    else:
        t = np.clip((k - warmup_steps) / (max_steps - warmup_steps), 0, 1)
        lr = np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t)
    return float(lr_init * (lr / lr_init))


def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, lr: float,
              eps: float = 1e-15, weight_decay: float = 0.0,
              betas: Tuple[float, float] = (0.9, 0.999)) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """One dense Adam update (step = 1-based count after the increment); returns new (p, m, v)."""
    b1, b2 = betas
    if weight_decay != 0:
        g = g + weight_decay * p
    m = torch.lerp(m, g, 1 - b1)
    v = v * b2 + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    step_size = lr / bc1
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - step_size * (m / denom)
    return p, m, v
