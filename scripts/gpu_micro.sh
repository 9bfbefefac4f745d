#!/bin/bash
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python scripts/gemm_microbench.py > gpurun_out/micro_fp16.json 2>&1
EMPOSE_TC_DEBUG=2 timeout -s KILL 300 python scripts/gemm_microbench.py > gpurun_out/micro_fp16_mmaonly.json 2>&1
EMPOSE_TC_DEBUG=4 timeout -s KILL 300 python scripts/gemm_microbench.py > gpurun_out/micro_fp16_noepi.json 2>&1
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "golden or oracle" > gpurun_out/pytest_quick.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_quick.log
tail -n 3 gpurun_out/pytest_quick.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -n 1 gpurun_out/bench.log | cut -c1-200
