set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "gemm_engine and fp32" > gpurun_out/t1_gemm_fp32.log 2>&1; echo "rc=$?" >> gpurun_out/t1_gemm_fp32.log
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "gemm_engine and tf32" > gpurun_out/t2_gemm_tf32.log 2>&1; echo "rc=$?" >> gpurun_out/t2_gemm_tf32.log
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "sensor_projection" > gpurun_out/t3_sensor.log 2>&1; echo "rc=$?" >> gpurun_out/t3_sensor.log
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ief and fp32" > gpurun_out/t4_ief_fp32.log 2>&1; echo "rc=$?" >> gpurun_out/t4_ief_fp32.log
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ief and tf32" > gpurun_out/t5_ief_tf32.log 2>&1; echo "rc=$?" >> gpurun_out/t5_ief_tf32.log
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "shards or host_buffer" > gpurun_out/t6_misc.log 2>&1; echo "rc=$?" >> gpurun_out/t6_misc.log
tail -5 gpurun_out/t*.log
