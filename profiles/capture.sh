#!/bin/bash
# ncu captures for profiles/ (run on the GPU box through gpurun:  gpurun -- 'bash profiles/capture.sh').
# One short EAGER bench per kernel family (the graph replays exactly these launches); raw-page CSVs come back,
# the large .ncu-rep files stay on the box (gpurun_out is capped at 64 MiB).
set -u
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 1 --eager --no-cpu-baseline --no-optimizer-leg"
# 1. launch list of two whole steps (+ warm-up): per-launch durations, cold-cache and serialised
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/r01_launches.csv $BENCH > $OUT/r01_launches.log 2>&1
cap() {  # cap <name> <kernel regex> <skip> <count>
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s "$3" -c "$4" -o $OUT/$1 $BENCH > $OUT/$1.log 2>&1
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1.raw.csv 2>/dev/null
  rm -f $OUT/$1.ncu-rep
}
cap r01_hash_bwd "hash_bwd" 16 4
cap r01_hash_fwd "hash_fwd" 16 4
cap r01_prop "prop_fwd_kernel|prop_bwd_kernel" 12 8
cap r01_mlp_tc_fwd "mlp_tc_fwd" 6 3
cap r01_mlp_tc_bwd "mlp_tc_bwd" 6 3
cap r01_ray "weights_fwd|weights_bwd|render_fwd|render_bwd|pdf_sample|piecewise" 30 10
cap r01_glue "field_split|density_act|distortion|interlevel|pixel_losses|density_l1|camera_opt|sh4|sample_pos" 40 16
# 2. the optimiser launch alone
ncu --set full --clock-control none -k regex:adam_kernel -s 3 -c 2 -o $OUT/r01_adam python tools/bench_adam.py > $OUT/r01_adam.log 2>&1
ncu -i $OUT/r01_adam.ncu-rep --page raw --csv > $OUT/r01_adam.raw.csv 2>/dev/null
rm -f $OUT/r01_adam.ncu-rep
ls -la $OUT
