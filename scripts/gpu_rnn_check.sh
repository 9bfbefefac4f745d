#!/bin/bash
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/parity_report.jsonl
timeout -s KILL 900 python -m pytest tests/test_gpu_rnn.py -q -m gpu > gpurun_out/pytest_rnn.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_rnn.log
tail -n 30 gpurun_out/pytest_rnn.log
cp gpurun_out/parity_report.jsonl gpurun_out/rnn_report.jsonl 2>/dev/null
timeout -s KILL 900 python bench.py --workload birnn --steps 5 --warmup 3 > gpurun_out/bench_birnn.log 2>&1; echo "rc=$?" >> gpurun_out/bench_birnn.log
tail -n 3 gpurun_out/bench_birnn.log | cut -c1-1500
