#!/bin/bash
# A/B of an environment switch on the 2-GPU bench, alternating runs on the same box
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$2', 'ms/step', round(d['ms_per_step'],4))"; }
for i in 1 2; do
  run 2951$i base
  export $1=1; run 2952$i "$1"; unset $1
done
