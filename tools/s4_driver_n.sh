N=$1; OUT=gpurun_out
P=$((29500 + RANDOM % 1000))
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 30 --warmup 5 > $OUT/s4_driver_n$N.json 2> $OUT/s4_driver_n$N.err ) 2>&1 | grep real
python - <<PY
import json
d=json.loads(open('$OUT/s4_driver_n$N.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('n_gpus','value','ms_per_step','scaling')}, 'e2e', d['e2e']['ms_per_step'], 'opt', d.get('with_optimizer',{}).get('ms_per_step'), d['config'].get('parallelism'), d['config'].get('exchange'))
PY
P=$((29500 + RANDOM % 1000))
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --impl reference --gpus $N --steps 1 --warmup 0 > $OUT/s4_driver_ref_n$N.json 2> $OUT/s4_driver_ref_n$N.err ) 2>&1 | grep real
tail -c 300 $OUT/s4_driver_ref_n$N.json
