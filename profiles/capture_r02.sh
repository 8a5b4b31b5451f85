#!/bin/bash
# Round-2 ncu captures (gpurun -- 'bash profiles/capture_r02.sh'): launch list of whole steps + one `--set full`
# capture per kernel family with raw and source pages.  One EAGER bench step (the graph replays exactly these launches).
set -u
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 1 --eager --no-cpu-baseline --no-optimizer-leg --train-only"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/r02g_launches.csv $BENCH > $OUT/r02g_launches.log 2>&1
cap() {  # cap <name> <kernel regex> <skip> <count> [source]
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s "$3" -c "$4" -o $OUT/$1 $BENCH > $OUT/$1.log 2>&1
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1.raw.csv 2>/dev/null
  if [ "${5:-}" = "source" ]; then ncu -i $OUT/$1.ncu-rep --page source --csv --print-source sass > $OUT/$1.source.csv 2>/dev/null; fi
  rm -f $OUT/$1.ncu-rep
}
cap r02g_hash_bwd "hash_bwd" 16 4 source
cap r02g_hash_fwd "hash_fwd" 16 4
cap r02g_prop "prop_fwd_kernel|prop_bwd_kernel" 12 8 source
cap r02g_mlp_tc_fwd "mlp_tc_fwd" 12 6
cap r02g_mlp_tc_bwd "mlp_tc_bwd" 12 6 source
cap r02g_level "level_resample|ray_heads|embed_bwd|ray_features" 20 12
ls -la $OUT | head -40
