#!/bin/bash
# Re-capture of the launch list and the proposal kernels after the zero-gradient early-out (same recipe as capture_r02.sh)
set -u
OUT=gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --eager --no-cpu-baseline --no-optimizer-leg --train-only"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/r02g_launches.csv $BENCH > $OUT/r02g_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"prop_fwd_kernel|prop_bwd_kernel" -s 12 -c 8 -o $OUT/r02g_prop $BENCH > $OUT/r02g_prop.log 2>&1
ncu -i $OUT/r02g_prop.ncu-rep --page raw --csv > $OUT/r02g_prop.raw.csv 2>/dev/null
rm -f $OUT/r02g_prop.ncu-rep
