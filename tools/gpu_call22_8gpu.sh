#!/bin/bash
# final tree on 8 GPUs: parity (default merge = NVSwitch multicast from four GPUs on), then the driver's weak-scaling command
set -x
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29564 tests/mgpu_check.py > gpurun_out/r2_c22_mgpu_check_w8.log 2>&1
echo "mgpu w8 rc=$?" >> gpurun_out/r2_c22_mgpu_check_w8.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c22_bench_n8.json 2> gpurun_out/r2_c22_bench_n8.err
echo "bench n8 rc=$?" >> gpurun_out/r2_c22_bench_n8.err
echo done
