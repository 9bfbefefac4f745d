#!/bin/bash
# shorter iteration: engine tests, microbench, bench in both cluster modes (+ optional debug bit), parity suites
tag=${1:-iter}; dbg=${2:-0}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k test_gemm_engine > gpurun_out/${tag}_gemm.log 2>&1; rc=$?
echo "rc=$rc" >> gpurun_out/${tag}_gemm.log; tail -n 3 gpurun_out/${tag}_gemm.log
[ $rc -ne 0 ] && exit 0
for cfg in "0 0" "2 0" "0 $dbg"; do
  set -- $cfg
  EMPOSE_TC_CLUSTER=$1 EMPOSE_TC_DEBUG=$2 timeout -s KILL 200 python scripts/gemm_microbench.py 131072x512x512 131072x512x2048 4096x2048x1024 > gpurun_out/micro_${tag}_c$1_d$2.json 2>&1
  echo "cluster=$1 dbg=$2"; grep -h tflops gpurun_out/micro_${tag}_c$1_d$2.json | tr -d '\n'; echo
  EMPOSE_TC_CLUSTER=$1 EMPOSE_TC_DEBUG=$2 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_c$1_d$2.log 2>&1
  tail -n 1 gpurun_out/bench_${tag}_c$1_d$2.log | cut -c1-220
done
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rnn.py tests/test_gpu_train.py -q -m gpu > gpurun_out/${tag}_parity.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_parity.log
tail -n 4 gpurun_out/${tag}_parity.log
EMPOSE_TC_CLUSTER=2 timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/${tag}_parity_c2.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_parity_c2.log
tail -n 3 gpurun_out/${tag}_parity_c2.log
