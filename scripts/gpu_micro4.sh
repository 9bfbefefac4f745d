#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export EMPOSE_TC_VERBOSE=1
for c in 1 0; do for m in 0 4 24; do
  EMPOSE_TC_CLUSTER=$c EMPOSE_TC_DEBUG=$m timeout -s KILL 300 python scripts/gemm_microbench.py 131072x512x512 4096x2048x1024 131072x512x2048 > gpurun_out/micro4_c${c}_m$m.json 2>&1
done; done
