OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/s4_pytest.log 2>&1; tail -3 $OUT/s4_pytest.log
timeout 300 python tools/trace_step.py --steps 2 --out $OUT/s4_trace 2>&1 | tail -3
