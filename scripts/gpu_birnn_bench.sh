#!/bin/bash
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python bench.py --workload birnn --steps 3 --warmup 3 > gpurun_out/bench_birnn.log 2>&1; echo "rc=$?" >> gpurun_out/bench_birnn.log
tail -n 3 gpurun_out/bench_birnn.log | cut -c1-1800
