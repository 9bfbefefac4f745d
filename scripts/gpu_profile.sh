#!/bin/bash
# Launch list + ncu --set full captures of the named kernels (outputs under gpurun_out/).
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
for spec in "$@"; do   # name:regex:skip
  name=${spec%%:*}; rest=${spec#*:}; regex=${rest%%:*}; skip=${rest#*:}
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/prof_$name -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$name.log 2>&1
done
ls -la gpurun_out
