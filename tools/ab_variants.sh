#!/bin/bash
# A/B of prebuilt library variants (scratch/variants/*.so, built here with OM_NVCC_EXTRA) and of
# environment switches on one box; prints K1 time, step time, rest, early phase per run.
run() {
  python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-config5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1'.ljust(34), 'step %.4f ms  K1 %.4f ms  frac %.3f  rest %.4f ms  early %.4f ms/step  deferred %.0f' % (d['ms_per_step'], r['kernel_ms'], r['frac'], r['rest_of_step_ms'], d['early_phase']['ms_per_step'], d['deferred_vertices_per_step']))"
}
for v in ${VARIANTS:-base2 nofb_u2 nofb_u1 nofb_u1_m8 nofb_u2_m8 fb_u1}; do
  cp scratch/variants/$v.so optimesh_b200/liboptimesh_b200.so
  run $v; run $v
done
