#!/bin/bash
# gradient-exchange schedules at N GPUs (gpurun --gpus N -- 'bash tools/bench_modes.sh N [mode:ENV=val,ENV=val ...]')
N=${1:-2}
shift
OUT=gpurun_out
CONFIGS=("$@")
[ ${#CONFIGS[@]} -eq 0 ] && CONFIGS=(pipeline overlap after)
for cfg in "${CONFIGS[@]}"; do
  mode=${cfg%%:*}
  envs=""
  [[ "$cfg" == *:* ]] && envs=$(echo "${cfg#*:}" | tr ',' ' ')
  tag=$(echo "$cfg" | tr ':=,' '___')
  env TN_COMM=$mode $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
    bench.py --gpus $N --steps 30 --warmup 5 --train-only > $OUT/modes_n${N}_$tag.json 2> $OUT/modes_n${N}_$tag.err
  python tools/show_modes.py | grep "n${N}_$tag.json"
done
