#!/bin/bash
# fast activations + cell-state prefetch in the LSTM epilogue, CTA-scope fence by default
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rnn.py tests/test_gpu_train.py -x -q -m gpu > gpurun_out/lstm_epi_parity.log 2>&1; echo "rc=$?" >> gpurun_out/lstm_epi_parity.log; tail -n 4 gpurun_out/lstm_epi_parity.log
for dbg in 0 128 64; do
  EMPOSE_TC_DEBUG=$dbg timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d${dbg}.log 2>&1
  echo "dbg=$dbg"; tail -n 1 gpurun_out/bench_d${dbg}.log | cut -c1-220
done
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout -s KILL 600 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train.log 2>&1; tail -n 1 gpurun_out/bench_train.log | cut -c1-220
