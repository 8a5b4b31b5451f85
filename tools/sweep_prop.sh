#!/bin/bash
OUT=gpurun_out
B="python bench.py --steps 10 --warmup 3 --train-only --no-cpu-baseline --no-optimizer-leg --profile-kernels"
run() {
  name=$1; shift
  env "$@" $B > $OUT/sp_$name.json 2> $OUT/sp_$name.err
  echo "== $name $*: $(python -c "import json;d=json.load(open('$OUT/sp_$name.json'));print(round(d['ms_per_step'],3))") ms/step"
  grep -E "prop_density_bwd|hash_encode_bwd|sum of" $OUT/sp_$name.err
}
run pr0 TN_PAIR_RED=0
run pr1 TN_PAIR_RED=1
run pr1a130 TN_PAIR_RED=1 TN_AGG_PROP=130
run pr1a300 TN_PAIR_RED=1 TN_AGG_PROP=300
run pr1a40 TN_PAIR_RED=1 TN_AGG_PROP=40
run pr1a0 TN_PAIR_RED=1 TN_AGG_PROP=0
