#!/bin/bash
# ncu captures of the encode kernels (gpurun -- 'bash profiles/capture_enc.sh <tag>'); raw + source pages come back.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 1 --eager --no-cpu-baseline --no-optimizer-leg --train-only"
cap() {  # cap <name> <kernel regex> <skip> <count>
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s "$3" -c "$4" -o $OUT/$1 $BENCH > $OUT/$1.log 2>&1
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1.raw.csv 2>/dev/null
  ncu -i $OUT/$1.ncu-rep --page source --csv --print-source sass > $OUT/$1.source.csv 2>/dev/null
  rm -f $OUT/$1.ncu-rep
}
cap ${TAG}_hash_bwd "hash_bwd" 16 2
cap ${TAG}_hash_fwd "hash_fwd" 16 2
