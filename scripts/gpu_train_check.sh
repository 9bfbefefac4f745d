#!/bin/bash
# GPU visit for the training path: training parity tests (all, no -x, so one visit shows every mismatch), then the rest.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/parity_report.jsonl
timeout -s KILL 900 python -m pytest tests/test_gpu_train.py -q -m gpu > gpurun_out/pytest_train.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_train.log
tail -n 40 gpurun_out/pytest_train.log
cp gpurun_out/parity_report.jsonl gpurun_out/train_report.jsonl 2>/dev/null
if [ "$1" = "all" ]; then
  timeout -s KILL 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_train.py > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
  tail -n 8 gpurun_out/pytest_gpu.log
  timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
  tail -n 3 gpurun_out/bench.log
fi
if [ "$1" = "bench" ] || [ "$2" = "bench" ]; then
  timeout -s KILL 600 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/bench_train.log 2>&1; echo "rc=$?" >> gpurun_out/bench_train.log
  tail -n 3 gpurun_out/bench_train.log
fi
