"""torch.profiler breakdown of one eager train step (which torch glue ops are left around the kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
import nerfstudio_thermal_b200 as tn
from nerfstudio_thermal_b200 import engine

args = bench.parse()
dev = torch.device("cuda")
model = bench.build_model(args).to(dev).train()
batch = {k: v.to(dev) for k, v in bench.make_batch(args.rays, 42).items()}
runner = engine.GraphedTrainStep(model, batch, use_graph=False)
for _ in range(3):
    runner.step(None)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    runner.step(None)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=50,
                                                         max_shapes_column_width=60))
