#!/bin/bash
# A/B of the working tree against an older commit checked out (and built) in scratch/old
run() { (cd $1 && timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'ms/step', round(d['ms_per_step'],4))"); }
for i in 1 2; do
  run scratch/old 2953$i
  run . 2954$i
done
