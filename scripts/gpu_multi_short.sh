#!/bin/bash
# 2-GPU check of the inference and training workloads under torchrun (short form)
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_infer_n$N.log
tail -n 2 gpurun_out/bench_infer_n$N.log | cut -c1-300
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 10 --warmup 3 --workload train --no-cpu-baseline > gpurun_out/bench_train_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_train_n$N.log
tail -n 2 gpurun_out/bench_train_n$N.log | cut -c1-300
