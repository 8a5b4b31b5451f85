mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1; tail -3 gpurun_out/s3_pytest.log
python tools/profile_glue.py > gpurun_out/s3_glue_ops.txt 2>&1
python bench.py --steps 20 --warmup 5 --train-only --no-cpu-baseline --no-optimizer-leg --profile-kernels > gpurun_out/s3_bench_base.json 2> gpurun_out/s3_bench_base.err
grep -E "launches/step|sum of" gpurun_out/s3_bench_base.err | head -40
