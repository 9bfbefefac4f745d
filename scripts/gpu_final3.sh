#!/bin/bash
# end-of-round verification (short form): every -m gpu test, smoke, bench (A/B against the per-item phases), train bench, launch list
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/parity_report.jsonl
timeout -s KILL 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
tail -n 2 gpurun_out/smoke.log
timeout -s KILL 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
tail -n 2 gpurun_out/bench.log | cut -c1-200
EMPOSE_MAIN_LEGACY_BLEND=1 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_legacy.log 2>&1
tail -n 1 gpurun_out/bench_legacy.log | cut -c1-200
timeout -s KILL 600 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train.log 2>&1
tail -n 1 gpurun_out/bench_train.log | cut -c1-200
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
