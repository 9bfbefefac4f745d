#!/bin/bash
# (6 frames per CTA, 4 CTAs per SM) variant of main_kernel on the slimmer frame state
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v54.log 2>&1; tail -n 1 gpurun_out/bench_v54.log | cut -c1-200
EMPOSE_MAIN_VARIANT=5 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v64.log 2>&1; tail -n 1 gpurun_out/bench_v64.log | cut -c1-200
EMPOSE_MAIN_VARIANT=5 timeout -s KILL 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_v64.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_v64.log; tail -n 3 gpurun_out/pytest_v64.log
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py tests/test_gpu_synth.py -q -m gpu > gpurun_out/pytest_v54.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_v54.log; tail -n 3 gpurun_out/pytest_v54.log
