"""Reading and writing the reference trainer's checkpoints.

Format of `Trainer.save_checkpoint` (engine/trainer.py:424-445): {"step", "pipeline": pipeline.state_dict() (model
keys prefixed "_model.", plus "module." under DDP), "optimizers": {group: torch.optim.Adam.state_dict()},
"schedulers": {group: LambdaLR.state_dict()}, "scalers": GradScaler.state_dict()}.  Loading follows
`Trainer._load_checkpoint` (:389-421) -> `VanillaPipeline.load_pipeline` (pipelines/base_pipeline.py:408-419) ->
`Pipeline.load_state_dict` (:100-125): prefixes stripped, model loaded with strict=True, the model told the step.
"""
from typing import Any, Dict, Optional

import torch

from .model import ThermalNerfactoModel
from .optim import FusedAdam, scheduled_lr


def model_state_from_pipeline(pipeline_state: Dict[str, Any]) -> Dict[str, Any]:
    """Pipeline.load_state_dict's key handling (base_pipeline.py:100-113)."""
    state = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in pipeline_state.items()}
    ddp, model_state = True, {}
    for k, v in state.items():
        if k.startswith("_model."):
            model_state[k[len("_model."):]] = v
            if not k.startswith("_model.module."):
                ddp = False
    if ddp:
        model_state = {k[len("module."):]: v for k, v in model_state.items()}
    return model_state


# Sub-modules of the reference model that hold weights but are reporting code, not the hot path: the torchmetrics
# objects built in NerfactoModel.populate_modules (models/nerfacto.py:251-253).  A real `step-*.ckpt` carries the
# LPIPS network's weights under `_model.lpips.net.*`; this package has no such modules, so those keys are dropped
# before the strict load instead of being reported as unexpected.
METRIC_MODULE_PREFIXES = ("lpips.", "psnr.", "ssim.")


def load_checkpoint(ckpt: Dict[str, Any], model: ThermalNerfactoModel, optimizer: Optional[FusedAdam] = None,
                    strict: Optional[bool] = True) -> int:
    """Load a reference `step-*.ckpt` dict (torch.load(..., map_location="cpu")) into the model (and the fused
    optimiser).  Returns the step to resume from (`loaded_state["step"] + 1`, trainer.py:402).

    Keys of the reference's metric modules (`METRIC_MODULE_PREFIXES`) are ignored.  Everything else is loaded with
    strict=True; `strict=None`/False retries non-strictly like Pipeline.load_state_dict (base_pipeline.py:116-122)."""
    step = int(ckpt["step"])
    model.set_anneal_step(step)  # Model.update_to_step
    state = {k: v for k, v in model_state_from_pipeline(ckpt["pipeline"]).items()
             if not k.startswith(METRIC_MODULE_PREFIXES)}
    try:
        model.load_state_dict(state, strict=True)
    except RuntimeError:
        if strict:
            raise
        model.load_state_dict(state, strict=False)
    if optimizer is not None and "optimizers" in ckpt:
        optimizer.load_state_dict({k: v for k, v in ckpt["optimizers"].items() if k in optimizer.config},
                                  schedulers=ckpt.get("schedulers"))
    return step + 1


def save_checkpoint(step: int, model: ThermalNerfactoModel, optimizer: Optional[FusedAdam] = None) -> Dict[str, Any]:
    """The dict `Trainer.save_checkpoint` writes (pass it to torch.save).  The reference's metric modules
    (`_model.lpips.net.*`) are not part of this package, so their keys are absent: the reference's trainer loads the
    result through its strict-then-non-strict fallback (Pipeline.load_state_dict with strict=None/False)."""
    out: Dict[str, Any] = {"step": step,
                           "pipeline": {f"_model.{k}": v.detach().cpu().clone() for k, v in model.state_dict().items()},
                           "optimizers": {}, "schedulers": {}, "scalers": {}}
    if optimizer is not None:
        out["optimizers"] = {k: _cpu(v) for k, v in optimizer.state_dict().items()}
        k = optimizer.step_count
        for name, cfg in optimizer.config.items():  # LambdaLR.state_dict() minus the (unpicklable) lambdas
            out["schedulers"][name] = {"base_lrs": [cfg.lr], "last_epoch": k, "_step_count": k + 1,
                                       "_last_lr": [scheduled_lr(cfg, k)], "lr_lambdas": [None]}
    return out


def _cpu(obj):
    if torch.is_tensor(obj):
        return obj.detach().cpu()
    if isinstance(obj, dict):
        return {k: _cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_cpu(v) for v in obj)
    return obj
