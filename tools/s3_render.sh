OUT=gpurun_out
python bench.py --mode render --steps 3 --warmup 1 --profile-kernels > $OUT/s3_render.json 2> $OUT/s3_render.err
grep -E "launches/step|sum of" $OUT/s3_render.err | head -30
python - <<'P'
import json
d=json.loads(open('gpurun_out/s3_render.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
for r in d.get('rooflines',[]): print({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items()})
P
