#!/bin/bash
one() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('[$1]', 'ms/step', round(d['ms_per_step'],4), 'K1', round(r['kernel_ms'],4), d['flips_in_timed_region'], d['limited_vertex_steps'])"; }
one base
for v in "-DOM_K1_BLOCK=128 -DOM_K1_MINB=8" "-DOM_K1_BLOCK=64 -DOM_K1_MINB=16"; do
  OM_NVCC_EXTRA="$v" python -m optimesh_b200.build 2>&1 | grep -i error
  OM_NVCC_EXTRA="$v" one "$v"
done
