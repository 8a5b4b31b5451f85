# usage: bash tools/s3_run.sh NAME "pytest args" "grep pattern"
OUT=gpurun_out; NAME=$1
if [ -n "$2" ]; then timeout 600 python -m pytest $2 -m gpu -x -q 2>&1 | tail -4; fi
B="python bench.py --steps 20 --warmup 5 --train-only --no-cpu-baseline --no-optimizer-leg --profile-kernels"
$B > $OUT/s3_$NAME.json 2> $OUT/s3_$NAME.err
echo "== $NAME: $(python -c "import json;d=json.load(open('$OUT/s3_$NAME.json'));print(round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))") ms/step"
grep -E "$3|sum of" $OUT/s3_$NAME.err
