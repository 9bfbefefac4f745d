#!/bin/bash
# e2e sweep of the pipelined host entry point: sub-batch sizes (EMPOSE_HOST_CHUNK) at the bench workload.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "host or pipelined" > gpurun_out/pytest_host.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_host.log
tail -n 5 gpurun_out/pytest_host.log
for c in 8192 2048 1024; do
  EMPOSE_HOST_CHUNK=$c timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chunk$c.log 2>&1
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_chunk$c.log').read().strip().splitlines()[-1])
print('chunk $c: device %.2f ms, e2e %.2f ms -> %.2fM frames/s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value'] / 1e6))
PY
done
