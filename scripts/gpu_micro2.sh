#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for m in 0 8 16 24 32 56; do
  EMPOSE_TC_DEBUG=$m timeout -s KILL 300 python scripts/gemm_microbench.py 131072x512x512 4096x2048x1024 18944x256x64 > gpurun_out/micro2_$m.json 2>&1
done
