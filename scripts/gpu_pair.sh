#!/bin/bash
# cta_group::2 variant of the executor (EMPOSE_TC_CLUSTER=2): correctness first (short timeouts: a protocol bug traps), then speed
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
EMPOSE_TC_CLUSTER=2 EMPOSE_TC_VERBOSE=1 timeout -s KILL 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k test_gemm_engine > gpurun_out/pair_gemm.log 2>&1; rc=$?
echo "rc=$rc" >> gpurun_out/pair_gemm.log; tail -n 12 gpurun_out/pair_gemm.log
if [ $rc -ne 0 ]; then
  timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k test_gemm_engine > gpurun_out/single_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/single_gemm.log; tail -n 3 gpurun_out/single_gemm.log
  timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_single.log 2>&1; tail -n 1 gpurun_out/bench_single.log | cut -c1-200
  exit 0
fi
for mode in 0 2; do
  EMPOSE_TC_CLUSTER=$mode timeout -s KILL 200 python scripts/gemm_microbench.py 131072x512x512 4096x2048x1024 131072x512x2048 4096x2048x672 > gpurun_out/micro_cluster$mode.json 2>&1
  EMPOSE_TC_CLUSTER=$mode EMPOSE_TC_DEBUG=4 timeout -s KILL 200 python scripts/gemm_microbench.py 131072x512x512 > gpurun_out/micro_cluster${mode}_noepi.json 2>&1
  EMPOSE_TC_CLUSTER=$mode timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cluster$mode.log 2>&1
  tail -n 1 gpurun_out/bench_cluster$mode.log | cut -c1-220
done
grep -h tflops gpurun_out/micro_cluster*.json
EMPOSE_TC_CLUSTER=2 timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rnn.py tests/test_gpu_train.py -q -m gpu > gpurun_out/pair_parity.log 2>&1; echo "rc=$?" >> gpurun_out/pair_parity.log
tail -n 8 gpurun_out/pair_parity.log
