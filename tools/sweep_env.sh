#!/bin/bash
# A/B of the encode tuning knobs through the headline bench's per-kernel table (gpurun -- 'bash tools/sweep_env.sh')
OUT=gpurun_out
B="python bench.py --steps 10 --warmup 3 --train-only --no-cpu-baseline --no-optimizer-leg --profile-kernels"
run() {  # run <name> VAR=val ...
  name=$1; shift
  env "$@" $B > $OUT/sw_$name.json 2> $OUT/sw_$name.err
  echo "== $name $*: $(python -c "import json;d=json.load(open('$OUT/sw_$name.json'));print(round(d['ms_per_step'],3))") ms/step"
  grep -E "hash_encode|sum of" $OUT/sw_$name.err
}
run base TN_PAIR_ENC=1e9 TN_PAIR_RED=0
run red TN_PAIR_ENC=1e9 TN_PAIR_RED=1
run all TN_PAIR_ENC=0 TN_PAIR_RED=1
run p100 TN_PAIR_ENC=100 TN_PAIR_RED=1
run p150 TN_PAIR_ENC=150 TN_PAIR_RED=1
run p250 TN_PAIR_ENC=250 TN_PAIR_RED=1
run p400 TN_PAIR_ENC=400 TN_PAIR_RED=1
run p150a200 TN_PAIR_ENC=150 TN_PAIR_RED=1 TN_AGG_ENC_PATCH=200
run p150a450 TN_PAIR_ENC=150 TN_PAIR_RED=1 TN_AGG_ENC_PATCH=450
run p150a800 TN_PAIR_ENC=150 TN_PAIR_RED=1 TN_AGG_ENC_PATCH=800
