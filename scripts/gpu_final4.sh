#!/bin/bash
# last check of the committed defaults: parity file, smoke, bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -q -m gpu > gpurun_out/pytest_final4.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_final4.log; tail -n 3 gpurun_out/pytest_final4.log
timeout -s KILL 200 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log
timeout -s KILL 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log; tail -n 2 gpurun_out/bench.log | cut -c1-200
