#!/bin/bash
# Re-capture of the launch list and the tensor-core MLP backward after the zero-gradient tile skip (recipe of capture_r02.sh)
set -u
OUT=gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --eager --no-cpu-baseline --no-optimizer-leg --train-only"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/r02g_launches.csv $BENCH > $OUT/r02g_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mlp_tc_bwd" -s 12 -c 6 -o $OUT/r02g_mlp_tc_bwd $BENCH > $OUT/r02g_mlp_tc_bwd.log 2>&1
ncu -i $OUT/r02g_mlp_tc_bwd.ncu-rep --page raw --csv > $OUT/r02g_mlp_tc_bwd.raw.csv 2>/dev/null
rm -f $OUT/r02g_mlp_tc_bwd.ncu-rep
