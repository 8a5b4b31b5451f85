#!/bin/bash
# ncu capture of the tensor-core MLP kernels (run on the GPU box through gpurun); CSV pages come back, reps stay.
set -u
OUT=gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --eager --no-cpu-baseline"
for k in mlp_tc_fwd mlp_tc_bwd; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 3 -o $OUT/r01_$k $BENCH > $OUT/r01_$k.log 2>&1
  ncu -i $OUT/r01_$k.ncu-rep --page raw --csv > $OUT/r01_$k.raw.csv 2>/dev/null
  ncu -i $OUT/r01_$k.ncu-rep --page source --csv > $OUT/r01_$k.source.csv 2>/dev/null
  rm -f $OUT/r01_$k.ncu-rep
done
ls -la $OUT | tail -8
