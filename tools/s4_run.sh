# usage: bash tools/s4_run.sh NAME  -> GPU tests, train-only bench with kernel profile, glue op attribution
OUT=gpurun_out; NAME=$1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/s4_${NAME}_pytest.log 2>&1; tail -15 $OUT/s4_${NAME}_pytest.log
B="python bench.py --steps 20 --warmup 5 --train-only --no-cpu-baseline --no-optimizer-leg --profile-kernels"
timeout 600 $B > $OUT/s4_$NAME.json 2> $OUT/s4_$NAME.err
echo "== $NAME: $(python -c "import json;d=json.load(open('$OUT/s4_$NAME.json'));print(round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'kernel ms', round(d['kernel_ms_per_step'],4))") ms/step"
tail -5 $OUT/s4_$NAME.err
timeout 300 python tools/profile_glue.py --train-only > $OUT/s4_${NAME}_glue.txt 2>&1; grep -v Warn $OUT/s4_${NAME}_glue.txt | tail -45
