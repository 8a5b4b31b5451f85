# usage: bash tools/s4_ab.sh "ENV=val ..." NAME  (train-only bench with per-kernel profile)
OUT=gpurun_out
B="python bench.py --steps 30 --warmup 5 --train-only --no-cpu-baseline --no-optimizer-leg"
for spec in "$@"; do
  name=$(echo "$spec" | tr ' =' '__')
  env $spec timeout 600 $B > $OUT/s4_ab_$name.json 2> $OUT/s4_ab_$name.err
  echo "== $spec: $(python -c "import json;d=json.load(open('$OUT/s4_ab_$name.json'));print(round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))" 2>&1 | tail -1) ms/step"
done
