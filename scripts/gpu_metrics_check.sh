#!/bin/bash
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/parity_report.jsonl
timeout -s KILL 600 python -m pytest tests/test_gpu_metrics.py -q -m gpu > gpurun_out/pytest_metrics.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_metrics.log
tail -n 30 gpurun_out/pytest_metrics.log
cp gpurun_out/parity_report.jsonl gpurun_out/metrics_report.jsonl 2>/dev/null
