#!/bin/bash
# ncu captures for profiles/ (run on the GPU box through gpurun).  One short eager bench per kernel family;
# raw-page CSVs come back, the (large) .ncu-rep files stay on the box except the dominant kernel's.
set -u
OUT=gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --eager --no-cpu-baseline"
cap() {  # cap <name> <kernel regex> <skip> <count>
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s "$3" -c "$4" -o $OUT/$1 $BENCH > $OUT/$1.log 2>&1
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1.raw.csv 2>/dev/null
}
cap r01_hash_bwd "hash_bwd" 16 4
ncu -i $OUT/r01_hash_bwd.ncu-rep --page source --csv --kernel-name regex:hash_bwd > $OUT/r01_hash_bwd.source.csv 2>/dev/null
cap r01_hash_fwd "hash_fwd" 16 4
cap r01_mlp_tc_bwd "mlp_tc_bwd" 10 5
cap r01_mlp_tc_fwd "mlp_tc_fwd" 10 5
cap r01_ray "weights_fwd|weights_bwd|render_fwd|render_bwd|pdf_sample" 30 8
rm -f $OUT/r01_hash_fwd.ncu-rep $OUT/r01_mlp_tc_bwd.ncu-rep $OUT/r01_mlp_tc_fwd.ncu-rep $OUT/r01_ray.ncu-rep
ls -la $OUT
