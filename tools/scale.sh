#!/bin/bash
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_r01_n$n.json 2> gpurun_out/scale_err_n$n.log
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_r01_n$n.json 2> gpurun_out/scale_err_n$n.log
  fi
  grep -E "rank 0\]|Error|error" gpurun_out/scale_err_n$n.log | cut -c1-250 | head -4
  python -c "
import json; d=json.load(open('gpurun_out/scale_r01_n$n.json')); print('N=$n', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['n_vertices'], 'K1', round(d['roofline']['kernel_ms'],3))"
done
