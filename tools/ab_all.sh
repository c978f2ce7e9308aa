#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run2() { (cd $1 && timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=2 $1', 'ms/step', round(d['ms_per_step'],4))"); }
run1() { (cd $1 && python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('N=1 $1', 'ms/step', round(d['ms_per_step'],4), 'K1', round(r['kernel_ms'],4), 'flip', round(r['flip_pass_ms'],4))"); }
run2 scratch/old 29531
run2 . 29541
run2 scratch/old 29532
run2 . 29542
run1 scratch/old
run1 .
