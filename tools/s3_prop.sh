OUT=gpurun_out
python -m pytest tests/test_gpu_fused.py tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 5 --train-only --no-cpu-baseline --no-optimizer-leg --profile-kernels"
for a in 96 130 300; do
  TN_AGG_PROP=$a $B > $OUT/s3_prop_a$a.json 2> $OUT/s3_prop_a$a.err
  echo "== TN_AGG_PROP=$a: $(python -c "import json;d=json.load(open('$OUT/s3_prop_a$a.json'));print(round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))") ms/step"
  grep -E "prop_density|sum of" $OUT/s3_prop_a$a.err
done
