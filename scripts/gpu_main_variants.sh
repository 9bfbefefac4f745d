#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in 0 1 2 3 4; do
  EMPOSE_MAIN_VARIANT=$v timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mv$v.log 2>&1
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_mv$v.log').read().strip().splitlines()[-1])
print('variant $v: %.2f ms/step  %.2fM frames/s' % (d['ms_per_step'], d['value'] / 1e6))
PY
done
for v in 1 3; do
  EMPOSE_MAIN_VARIANT=$v timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "golden or oracle" > gpurun_out/pytest_mv$v.log 2>&1
  tail -n 1 gpurun_out/pytest_mv$v.log
done
