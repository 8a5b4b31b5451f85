OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_model.py tests/test_gpu_levels.py -m gpu -x -q 2>&1 | tail -2
B="python bench.py --steps 20 --warmup 5 --train-only --no-cpu-baseline --no-optimizer-leg --profile-kernels"
timeout 600 $B > $OUT/s4_prop.json 2> $OUT/s4_prop.err
python -c "import json;d=json.load(open('$OUT/s4_prop.json'));print(round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"
grep -E "prop_density|sum of" $OUT/s4_prop.err
python - <<'PY'
import torch, bench, sys
from nerfstudio_thermal_b200 import engine, fused_ops
args = bench.parse()
dev = torch.device("cuda")
model = bench.build_model(args).to(dev).train()
batch = {k: v.to(dev) for k, v in bench.make_batch(args.rays, 42).items()}
seen = []
orig = fused_ops._PropDensityFn.backward
def spy(ctx, d_density, *a):
    d = d_density.view(-1)
    n = d.numel() // 32 * 32
    w = (d[:n].view(-1, 32) != 0).any(dim=1).float().mean().item()
    seen.append((d.numel(), round((d != 0).float().mean().item(), 4), round(w, 4)))
    return orig(ctx, d_density, *a)
fused_ops._PropDensityFn.backward = staticmethod(spy)
runner = engine.GraphedTrainStep(model, batch, use_graph=False)
runner.step(None); torch.cuda.synchronize(); seen.clear()
runner.step(None); torch.cuda.synchronize()
print("prop backward calls: (samples, nonzero fraction of d_density, fraction of warps with any nonzero)", seen)
PY
