OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/s4_mlp_pytest.log 2>&1; tail -4 $OUT/s4_mlp_pytest.log
bash tools/s4_ab.sh TN_X=stack
timeout 600 python bench.py --mode render --steps 3 --warmup 1 --no-cpu-baseline > $OUT/s4_mlp_render.json 2> $OUT/s4_mlp_render.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s4_mlp_render.json'))
print('render', round(d['value']), d['unit'], 'e2e', d.get('e2e',{}).get('value'))
PY
