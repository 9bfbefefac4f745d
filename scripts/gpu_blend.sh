#!/bin/bash
# CTA-wide vectorised shape-blend phases of main_kernel: parity, then A/B against the per-item phases
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synth.py tests/test_gpu_train.py -x -q -m gpu > gpurun_out/blend_parity.log 2>&1; echo "rc=$?" >> gpurun_out/blend_parity.log; tail -n 4 gpurun_out/blend_parity.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_blend_new.log 2>&1; tail -n 1 gpurun_out/bench_blend_new.log | cut -c1-200
EMPOSE_MAIN_LEGACY_BLEND=1 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_blend_old.log 2>&1; tail -n 1 gpurun_out/bench_blend_old.log | cut -c1-200
