"""2-rank smoke of the overlapped gradient all-reduce (run under torchrun with a hard timeout)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from nerfstudio_thermal_b200 import engine, parallel


def say(*a):
    print(f"[rank {os.environ.get('RANK')}] {time.strftime('%H:%M:%S')}", *a, flush=True)


rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
args = bench.parse()
model = bench.build_model(args).to(dev).train()
batch = {k: v.to(dev) for k, v in bench.make_batch(args.rays, parallel.rank_seed(42, rank)).items()}
use_graph = os.environ.get("TN_DEBUG_GRAPH", "1") == "1"
say("building runner, graph =", use_graph, "comm =", os.environ.get("TN_COMM", "overlap"))
runner = engine.GraphedTrainStep(model, batch, use_graph=use_graph)
say("runner built; early_end", runner._early_end, "comm_in_graph", runner._comm_in_graph)
for i in range(5):
    t = runner.step(None)
    torch.cuda.synchronize()
    say("step", i, float(t))
# gradients must be identical on both ranks after the averaged all-reduce
g = runner.grads.flat
ref = g.clone()
dist.broadcast(ref, 0)
say("max |grad - rank0 grad| =", float((g - ref).abs().max()), " |grad| =", float(g.abs().max()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier(); torch.cuda.synchronize()
e0.record()
for _ in range(20):
    runner.step(None)
e1.record(); torch.cuda.synchronize()
say("ms/step", e0.elapsed_time(e1) / 20)
sys.stdout.flush(); bench.shutdown_distributed()
