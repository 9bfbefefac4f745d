#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python scripts/gemm_microbench.py 131072x512x512 4096x2048x1024 18944x256x64 > gpurun_out/micro3.json 2>&1
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/pytest_quick.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_quick.log
tail -n 5 gpurun_out/pytest_quick.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -n 1 gpurun_out/bench.log | cut -c1-200
