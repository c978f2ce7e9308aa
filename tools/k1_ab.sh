#!/bin/bash
# A/B of K1 variants on one box: "ENV=..;NVCC=.." specs, each built (if NVCC flags differ) and
# run through the short bench; prints K1 kernel time, flip pass and step time.
#   tools/k1_ab.sh "base" "OM_K1_PREFETCH_BLOCKS=0" "NVCC:-DOM_K1_MINB=6"
run() {
  python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-config5 ${BENCH_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1'.ljust(44), 'step %.4f ms  K1 %.4f ms  frac %.3f  rest %.4f ms  early %.4f ms/step' % (d['ms_per_step'], r['kernel_ms'], r['frac'], r['rest_of_step_ms'], d['early_phase']['ms_per_step']))"
}
for spec in "$@"; do
  if [[ "$spec" == NVCC:* ]]; then
    OM_NVCC_EXTRA="${spec#NVCC:}" python -m optimesh_b200.build > /dev/null 2>&1 || echo "build failed: $spec"
    run "$spec"; run "$spec"
    python -m optimesh_b200.build > /dev/null 2>&1
  elif [[ "$spec" == base ]]; then
    run base; run base
  else
    env $spec bash -c "$(declare -f run); BENCH_ARGS='$BENCH_ARGS' run '$spec'; BENCH_ARGS='$BENCH_ARGS' run '$spec'"
  fi
done
