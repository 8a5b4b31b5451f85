"""Kernel timeline of a few train steps (torch profiler / CUPTI) on rank 0, as a compact table: stream, start, duration,
name -- to see what the gradient exchange overlaps with.   torchrun ... tools/trace_step.py [--steps 3]
TN_COMM selects the exchange schedule; TN_PIPELINE_FORCE=1 runs the three-phase schedule on one GPU."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nerfstudio_thermal_b200 import engine, parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/trace")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    D = bench.Dist(world)
    args = argparse.Namespace(density_mode="separate", log2_hashmap_size=19, init="trained")
    model = bench.build_model(args).to(D.dev).train()
    batch = {k: v.to(D.dev) for k, v in bench.make_batch(4096, parallel.rank_seed(42, D.rank)).items()}
    force = True if os.environ.get("TN_PIPELINE_FORCE") == "1" else None
    runner = engine.GraphedTrainStep(model, batch, use_graph=True, pipeline=force)
    for _ in range(10):
        runner.step(None)
    D.barrier()
    ms = D.timed(lambda: runner.step(None), 30) / 30
    if D.rank == 0:
        print(f"mode={os.environ.get('TN_COMM', 'default')} world={world} pipeline={runner._pipeline}: {ms:.3f} ms/step")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            runner.step(None)
        torch.cuda.synchronize()
    D.barrier()
    if D.rank == 0:
        path = f"{a.out}_n{world}_{os.environ.get('TN_COMM', 'default')}.json"
        prof.export_chrome_trace(path + ".full")
        ev = [e for e in json.load(open(path + ".full"))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
        os.remove(path + ".full")
        t0 = min(e["ts"] for e in ev)
        rows = sorted(((e["ts"] - t0, e["dur"], e["args"].get("stream", -1), e["name"][:60]) for e in ev))
        json.dump(rows, open(path, "w"))
        print(f"wrote {path}: {len(rows)} GPU activities")
    if world > 1:
        bench.shutdown_distributed()


if __name__ == "__main__":
    main()
