#!/bin/bash
# end-of-round verification: main_kernel variant sweep, every -m gpu test, smoke, all bench workloads, launch list, ncu of main_kernel
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/parity_report.jsonl
for v in 0 2 3; do
  EMPOSE_MAIN_VARIANT=$v timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_variant$v.log 2>&1
  tail -n 1 gpurun_out/bench_variant$v.log | cut -c1-180
done
timeout -s KILL 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
tail -n 2 gpurun_out/smoke.log
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1
timeout -s KILL 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
tail -n 2 gpurun_out/bench.log | cut -c1-300
timeout -s KILL 600 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/bench_train.log 2>&1
timeout -s KILL 600 python bench.py --workload synth --steps 10 --warmup 3 > gpurun_out/bench_synth.log 2>&1
timeout -s KILL 600 python bench.py --workload birnn --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_birnn.log 2>&1
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:main_kernel -s 2 -c 1 -o gpurun_out/prof_main -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_main.log 2>&1
ls gpurun_out | head -50
