#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for spec in chain:gemm_tc:37 lstm:gemm_tc:12; do
  name=${spec%%:*}; rest=${spec#*:}; regex=${rest%%:*}; skip=${rest#*:}
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/prof2_$name -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2_$name.log 2>&1
done
EMPOSE_TC_DEBUG=2 timeout -s KILL 300 ncu --set full --clock-control none -k regex:gemm_tc -s 10 -c 1 -o gpurun_out/prof2_mmaonly -f \
      python scripts/gemm_microbench.py 131072x512x2048 > gpurun_out/ncu2_mmaonly.log 2>&1
EMPOSE_TC_DEBUG=4 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 10 -c 1 -o gpurun_out/prof2_noepi -f \
      python scripts/gemm_microbench.py 131072x512x512 > gpurun_out/ncu2_noepi.log 2>&1
ls -la gpurun_out/*.ncu-rep
