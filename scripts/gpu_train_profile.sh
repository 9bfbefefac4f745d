#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --workload train --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_train.log 2>&1
tail -2 gpurun_out/ncu_launches_train.log | cut -c1-200
