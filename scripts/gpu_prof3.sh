#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for spec in chain:37 lstm:12 tf32a:36 tf32b:35; do
  name=${spec%%:*}; skip=${spec#*:}
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s $skip -c 1 -o gpurun_out/prof3_$name -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3_$name.log 2>&1
done
ls -la gpurun_out/prof3*.ncu-rep
