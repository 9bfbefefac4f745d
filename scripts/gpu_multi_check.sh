#!/bin/bash
# N-GPU check of both bench workloads under torchrun (N = $1, default 2).
set -x
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_infer_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_infer_n$N.log
tail -n 2 gpurun_out/bench_infer_n$N.log | cut -c1-400
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 10 --warmup 3 --workload train > gpurun_out/bench_train_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_train_n$N.log
tail -n 2 gpurun_out/bench_train_n$N.log | cut -c1-600
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 2 --warmup 1 --impl reference > gpurun_out/bench_ref_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref_n$N.log
tail -n 2 gpurun_out/bench_ref_n$N.log | cut -c1-300
