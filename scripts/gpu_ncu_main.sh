#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 70 ncu --set full --clock-control none --import-source on -k regex:main_kernel -s 2 -c 1 -o gpurun_out/prof_main_v14 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_main_v14.log 2>&1
ls -la gpurun_out/prof_main_v14.ncu-rep
