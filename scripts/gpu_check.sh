#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, launch list and ncu captures (outputs under gpurun_out/).
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 15 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
tail -n 3 gpurun_out/smoke.log
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
tail -n 4 gpurun_out/bench.log
if [ "$1" = "profile" ]; then
  timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 37 -c 1 -o gpurun_out/prof_chain -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_chain.log 2>&1
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 12 -c 1 -o gpurun_out/prof_lstm -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_lstm.log 2>&1
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:main_kernel -s 2 -c 1 -o gpurun_out/prof_main -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_main.log 2>&1
  ls -la gpurun_out
fi
