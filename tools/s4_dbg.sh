OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -x -q > $OUT/s4_dbg_engine.log 2>&1; tail -5 $OUT/s4_dbg_engine.log
