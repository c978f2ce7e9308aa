#!/bin/bash
for v in "" "-DOM_SUSPECT_PREFETCH_ADJ" "-DOM_K1_CPASYNC_CG" "-DOM_K1_BLOCK=128 -DOM_K1_MINB=8" "-DOM_K1_UNROLL=3"; do
  OM_NVCC_EXTRA="$v" python -m optimesh_b200.build 2>&1 | grep -i error
  OM_NVCC_EXTRA="$v" python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > /tmp/b.json 2>/tmp/b.err
  python -c "
import json; d=json.load(open('/tmp/b.json')); r=d['roofline']; print('[$v]', 'ms/step', round(d['ms_per_step'],4), 'K1', round(r['kernel_ms'],4), 'flip', round(r['flip_pass_ms'],4), d['flips_in_timed_region'])"
done
