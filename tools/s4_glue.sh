OUT=gpurun_out
python tools/profile_glue.py --train-only > $OUT/s4_glue_ops.txt 2>&1; tail -70 $OUT/s4_glue_ops.txt
