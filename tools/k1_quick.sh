#!/bin/bash
# quick look at the ring kernel: parity tests that touch it, the bench line, instruction counts
tag=${1:-k1q}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-config5 --no-e2e > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --rounds 0 --no-e2e --no-cpu-baseline --no-config5 > gpurun_out/${tag}_bench_grid.json 2>> gpurun_out/${tag}_bench.err
OM_NO_GRAPH=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:k_step_ring -c 3 --csv --log-file gpurun_out/${tag}_k1.csv python bench.py --steps 3 --warmup 17 --no-e2e --no-cpu-baseline --no-config5 > gpurun_out/${tag}_ncu.log 2>&1
grep -v "^==" gpurun_out/${tag}_k1.csv | cut -d, -f5,13- | tail -16
