#!/bin/bash
# four GPUs, the one world size not yet exercised with the final tree (merge = NVSwitch multicast from four GPUs on)
set -x
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_c30_bench_n4.json 2> gpurun_out/r2_c30_bench_n4.err
echo "bench n4 rc=$?" >> gpurun_out/r2_c30_bench_n4.err
echo done
