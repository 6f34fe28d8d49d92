#!/bin/bash
# 8 GPUs with the NVSwitch-multicast merge (KBBQ_MERGE=nvls): parity first, then the weak-scaling bench and the human-scale
# geometry; every command is bounded so that a stuck multicast bind cannot eat the budget
set -x
mkdir -p gpurun_out
KBBQ_MERGE=nvls KBBQ_FILTER_VMM=1 KBBQ_NVLS_VERBOSE=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29564 tests/mgpu_check.py > gpurun_out/r2_c9_mgpu_check_nvls_w8.log 2>&1
echo "mgpu nvls w8 rc=$?" >> gpurun_out/r2_c9_mgpu_check_nvls_w8.log
KBBQ_MERGE=nvls KBBQ_NVLS_VERBOSE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c9_bench_n8.json 2> gpurun_out/r2_c9_bench_n8.err
echo "bench n8 rc=$?" >> gpurun_out/r2_c9_bench_n8.err
KBBQ_MERGE=nvls KBBQ_NVLS_VERBOSE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --genome 387500000 --coverage 30 > gpurun_out/r2_c9_bench_human_n8.json 2> gpurun_out/r2_c9_bench_human_n8.err
echo "bench human rc=$?" >> gpurun_out/r2_c9_bench_human_n8.err
echo done
