#!/bin/bash
# 2 GPUs: multi-GPU parity (python/NCCL path and the C++ Session path), a short 2-GPU bench
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_c5_gpus.txt 2>&1
( time timeout 1500 python -m pytest tests/test_mgpu.py tests/test_cli_gpu.py -m gpu -q -k "sharded or mgpu or streamed" ) > gpurun_out/r2_c5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c5_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c5_bench_n2.json 2> gpurun_out/r2_c5_bench_n2.err
echo "bench rc=$?" >> gpurun_out/r2_c5_bench_n2.err
echo done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c5_bench_c3.json 2> gpurun_out/r2_c5_bench_c3.err
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "insert_modes or golden or known" > gpurun_out/r2_c5_pytest_modes.log 2>&1
echo done2
