"""Which python lines emit the torch (aten) ops of one eager train step: a TorchDispatchMode logs every op that
reaches the CUDA backend with the innermost frame inside this package (forward and custom-Function backwards;
formula backwards of built-in ops show up as "(autograd)")."""
import sys, os, collections, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.utils._python_dispatch import TorchDispatchMode
import bench
from nerfstudio_thermal_b200 import engine

SKIP = ("aten.view", "aten.reshape", "aten._unsafe_view", "aten.detach", "aten.alias", "aten.expand", "aten.slice",
        "aten.select", "aten.unsqueeze", "aten.squeeze", "aten.t.", "aten.transpose", "aten.as_strided", "aten.empty",
        "aten.permute", "aten.split", "aten.unbind", "aten.is_", "aten.sym_", "aten.size", "aten.stride",
        "aten.lift_fresh", "aten._local_scalar", "aten.new_empty", "aten.narrow", "aten.empty_like")


class Log(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.agg = collections.defaultdict(int)

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = str(func)
        if not name.startswith(SKIP):
            where = "(autograd)"
            for fr in reversed(traceback.extract_stack()):
                if "nerfstudio_thermal_b200" in fr.filename and "tools" not in fr.filename:
                    where = f"{os.path.basename(fr.filename)}:{fr.lineno} {fr.name}"
                    break
            shapes = [tuple(a.shape) for a in args if torch.is_tensor(a)][:3]
            self.agg[(name, str(shapes), where)] += 1
        return func(*args, **(kwargs or {}))


args = bench.parse()
dev = torch.device("cuda")
model = bench.build_model(args).to(dev).train()
batch = {k: v.to(dev) for k, v in bench.make_batch(args.rays, 42).items()}
runner = engine.GraphedTrainStep(model, batch, use_graph=False)
for _ in range(2):
    runner.step(None)
torch.cuda.synchronize()
log = Log()
with log:
    runner.step(None)
torch.cuda.synchronize()
print("ops:", sum(log.agg.values()))
bywhere = collections.defaultdict(int)
for (n, s, w), c in log.agg.items():
    bywhere[w] += c
for w, c in sorted(bywhere.items(), key=lambda kv: -kv[1])[:70]:
    print(f"{c:4d}  {w}")
print("---- (autograd) ops by name/shape")
for (n, s, w), c in sorted(log.agg.items(), key=lambda kv: -kv[1]):
    if w == "(autograd)":
        print(f"{c:4d}  {n:34s} {s}")
