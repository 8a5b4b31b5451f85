"""Micro-benchmark of the fused Adam launch on a flat buffer of the model's size (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerfstudio_thermal_b200 import optim
from nerfstudio_thermal_b200.parallel import FlatGradBuffer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 38_800_000
ps = {"fields": [torch.nn.Parameter(torch.randn(n // 2, device="cuda"))],
      "proposal_networks": [torch.nn.Parameter(torch.randn(n // 2, device="cuda"))]}
buf = FlatGradBuffer.from_param_groups(ps)
opt = optim.FusedAdam(buf, optim.thermal_nerfacto_optimizers())
buf.flat.normal_()
for zero in (False, True):
    for _ in range(3):
        opt.step(zero_grads=zero)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        opt.step(zero_grads=zero)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nbytes = n * 4 * (8 if zero else 7)
    print(f"zero_grads={zero}: {ms:.3f} ms/step, {nbytes / ms / 1e6:.0f} GB/s")
