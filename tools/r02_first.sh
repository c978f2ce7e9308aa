#!/bin/bash
# round 2, first GPU pass: parity suite, a short bench, launch list and one full capture of K1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02a_tests.txt
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_ring -s 2 -c 2 -o gpurun_out/r02a_kstep python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
tail -3 gpurun_out/r02a_tests.txt; cat gpurun_out/r02a_bench.json | head -c 1500
