#!/bin/bash
# single-GPU A/B of the working tree against scratch/old
run() { (cd $1 && python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1', 'ms/step', round(d['ms_per_step'],4), 'K1', round(r['kernel_ms'],4), 'flip', round(r['flip_pass_ms'],4))"); }
for i in 1 2; do
  run scratch/old
  run .
done
