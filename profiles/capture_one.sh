set -u
OUT=gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --eager --no-cpu-baseline --no-optimizer-leg"
ncu --set full --clock-control none --import-source on -k regex:mlp_tc_fwd -s 6 -c 2 -o $OUT/mlpf $BENCH > $OUT/mlpf.log 2>&1
ncu -i $OUT/mlpf.ncu-rep --page raw --csv > $OUT/mlpf.raw.csv 2>/dev/null
ncu -i $OUT/mlpf.ncu-rep --page source --csv > $OUT/mlpf.source.csv 2>/dev/null
rm -f $OUT/mlpf.ncu-rep
