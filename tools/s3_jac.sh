OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "jacobian or hash" 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 5 --train-only --no-cpu-baseline --no-optimizer-leg --profile-kernels"
run() { name=$1; shift
  env "$@" $B > $OUT/s3_jac_$name.json 2> $OUT/s3_jac_$name.err
  echo "== $name $*: $(python -c "import json;d=json.load(open('$OUT/s3_jac_$name.json'));print(round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))") ms/step"
  grep -E "hash_encode|sum of" $OUT/s3_jac_$name.err
}
run off TN_SAVE_JAC=0
run j150 TN_SAVE_JAC=1 TN_JAC_ENC=150
run j400 TN_SAVE_JAC=1 TN_JAC_ENC=400
run j50 TN_SAVE_JAC=1 TN_JAC_ENC=50
run j0 TN_SAVE_JAC=1 TN_JAC_ENC=0
run off2 TN_SAVE_JAC=0
