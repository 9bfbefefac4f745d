#!/bin/bash
# Full GPU visit: every -m gpu test, smoke, the default bench, the other workloads, launch list and ncu captures.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/parity_report.jsonl
timeout -s KILL 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
tail -n 2 gpurun_out/smoke.log
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1
timeout -s KILL 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
tail -n 2 gpurun_out/bench.log | cut -c1-300
timeout -s KILL 600 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/bench_train.log 2>&1
timeout -s KILL 600 python bench.py --workload synth --steps 10 --warmup 3 > gpurun_out/bench_synth.log 2>&1
tail -n 1 gpurun_out/bench_synth.log | cut -c1-400
timeout -s KILL 600 python bench.py --workload metrics --steps 10 --warmup 3 > gpurun_out/bench_metrics.log 2>&1
tail -n 1 gpurun_out/bench_metrics.log | cut -c1-400
timeout -s KILL 600 python bench.py --workload birnn --steps 5 --warmup 3 > gpurun_out/bench_birnn.log 2>&1
tail -n 1 gpurun_out/bench_birnn.log | cut -c1-400
if [ "$1" = "profile" ]; then
  timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
  for spec in chain:gemm_tc:37 lstm:gemm_tc:12 main:main_kernel:2; do
    name=${spec%%:*}; rest=${spec#*:}; regex=${rest%%:*}; skip=${rest#*:}
    timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/prof_$name -f \
        python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$name.log 2>&1
  done
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:lstm_persistent -s 0 -c 1 -o gpurun_out/prof_persistent -f \
      python bench.py --workload birnn --frames 4000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_persistent.log 2>&1
fi
ls -la gpurun_out | head -40
