OUT=gpurun_out
python -m pytest tests -m gpu -x -q > $OUT/s3_pytest_full.log 2>&1; tail -2 $OUT/s3_pytest_full.log
python bench.py > $OUT/s3_bench_full.json 2> $OUT/s3_bench_full.err; tail -c 600 $OUT/s3_bench_full.err
python bench.py --impl reference --steps 1 --warmup 0 > $OUT/s3_bench_ref.json 2> $OUT/s3_bench_ref.err
