#!/bin/bash
# job descriptors in shared memory / registers; CTA-scope fence experiment (EMPOSE_TC_DEBUG=64)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/fence_parity0.log 2>&1; echo "rc=$?" >> gpurun_out/fence_parity0.log; tail -n 3 gpurun_out/fence_parity0.log
for dbg in 0 64; do
  for mode in 0 2; do
    EMPOSE_TC_DEBUG=$dbg EMPOSE_TC_CLUSTER=$mode timeout -s KILL 200 python scripts/gemm_microbench.py 131072x512x512 4096x2048x1024 4096x2048x672 > gpurun_out/micro_d${dbg}_c$mode.json 2>&1
    EMPOSE_TC_DEBUG=$dbg EMPOSE_TC_CLUSTER=$mode timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d${dbg}_c$mode.log 2>&1
    echo "dbg=$dbg cluster=$mode"; tail -n 1 gpurun_out/bench_d${dbg}_c$mode.log | cut -c1-220
    grep -h tflops gpurun_out/micro_d${dbg}_c$mode.json | tr -d '\n'; echo
  done
done
EMPOSE_TC_DEBUG=64 timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rnn.py tests/test_gpu_train.py -q -m gpu > gpurun_out/fence_parity64.log 2>&1; echo "rc=$?" >> gpurun_out/fence_parity64.log
tail -n 4 gpurun_out/fence_parity64.log
EMPOSE_TC_DEBUG=64 EMPOSE_TC_CLUSTER=2 timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/fence_parity64_c2.log 2>&1; echo "rc=$?" >> gpurun_out/fence_parity64_c2.log
tail -n 4 gpurun_out/fence_parity64_c2.log
