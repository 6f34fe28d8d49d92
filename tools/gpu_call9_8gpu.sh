#!/bin/bash
# 8 GPUs with the NVSwitch-multicast merge: default weak-scaling bench, the human-scale geometry, multi-GPU parity at N=8 and N=4
set -x
mkdir -p gpurun_out
KBBQ_NVLS_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c9_bench_n8.json 2> gpurun_out/r2_c9_bench_n8.err
echo "bench n8 rc=$?" >> gpurun_out/r2_c9_bench_n8.err
KBBQ_NVLS_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --genome 387500000 --coverage 30 > gpurun_out/r2_c9_bench_human_n8.json 2> gpurun_out/r2_c9_bench_human_n8.err
echo "bench human rc=$?" >> gpurun_out/r2_c9_bench_human_n8.err
KBBQ_NVLS_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c9_bench_n4.json 2> gpurun_out/r2_c9_bench_n4.err
echo "bench n4 rc=$?" >> gpurun_out/r2_c9_bench_n4.err
KBBQ_MERGE=nvls KBBQ_FILTER_VMM=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29564 tests/mgpu_check.py > gpurun_out/r2_c9_mgpu_check_nvls_w8.log 2>&1
echo done
