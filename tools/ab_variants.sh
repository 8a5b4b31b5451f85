#!/bin/bash
# A/B of prebuilt library variants on the GPU box: tools/ab_variants.sh "grep-pattern" v1 v2 ...   ("base" = the default build)
OUT=gpurun_out; L=nerfstudio_thermal_b200/lib
PAT=$1; shift
cp $L/libtn_b200.so /tmp/base.so
B="python bench.py --steps 20 --warmup 5 --train-only --no-cpu-baseline --no-optimizer-leg --profile-kernels"
for v in "$@"; do
  if [ "$v" = base ]; then cp /tmp/base.so $L/libtn_b200.so; else cp $L/variants/$v.so $L/libtn_b200.so; fi
  $B > $OUT/ab_$v.json 2> $OUT/ab_$v.err
  echo "== $v: $(python -c "import json;d=json.load(open('$OUT/ab_$v.json'));print(round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))") ms/step"
  grep -E "$PAT|sum of" $OUT/ab_$v.err
done
cp /tmp/base.so $L/libtn_b200.so
