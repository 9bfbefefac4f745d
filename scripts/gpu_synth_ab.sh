#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python bench.py --workload synth --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/synth_new.log 2>&1; tail -n 1 gpurun_out/synth_new.log | cut -c1-260
EMPOSE_MAIN_LEGACY_BLEND=1 timeout -s KILL 300 python bench.py --workload synth --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/synth_old.log 2>&1; tail -n 1 gpurun_out/synth_old.log | cut -c1-260
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_synth.csv \
    python bench.py --workload synth --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_synth.log 2>&1
tail -n 14 gpurun_out/launches_synth.csv | cut -d, -f5,12- | cut -c1-150
