#!/bin/bash
# round 2: the bench line, the launch list of the same command and full captures of the top kernels
tag=${1:-r02b}
if [ -z "$ONLY_NCU" ]; then
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --rounds 0 --no-e2e --no-cpu-baseline --no-config5 > gpurun_out/${tag}_bench_grid.json 2>> gpurun_out/${tag}_bench.err
python bench.py --method lloyd --no-e2e --no-cpu-baseline --no-config5 > gpurun_out/${tag}_bench_lloyd.json 2>> gpurun_out/${tag}_bench.err
fi
OM_NO_GRAPH=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-config5 > gpurun_out/${tag}_ncu1.log 2>&1
OM_NO_GRAPH=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_step_ring -c 4 -o gpurun_out/${tag}_kstep python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-config5 > gpurun_out/${tag}_ncu2.log 2>&1
OM_NO_GRAPH=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_suspect_flags|k_post|k_walk_list|k_reduce_stats|k_flip1|k_flip2|k_build_rings" -c 14 -o gpurun_out/${tag}_rest python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-config5 > gpurun_out/${tag}_ncu3.log 2>&1
ls -la gpurun_out/ | grep ${tag}
if [ -z "$ONLY_NCU" ]; then head -c 3000 gpurun_out/${tag}_bench.json; echo; tail -5 gpurun_out/${tag}_bench.err; fi
