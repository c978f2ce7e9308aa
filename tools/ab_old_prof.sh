#!/bin/bash
run() { (cd $1 && OM_DIST_PROFILE=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | grep -E "^\[rank 0\]|^\{" | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$1', 'ms/step', round(d['ms_per_step'],4), {k:v for k,v in d.items() if k in ('band_vertices_all_ranks','fallback_full_gathers','slow_flip_rounds','flips_in_timed_region','flip_rounds_in_timed_region','gpu_launches')})
    else: print('$1', line.strip()[:900])"); }
run scratch/old 29531
run . 29541
