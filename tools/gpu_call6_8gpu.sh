#!/bin/bash
# 8 GPUs: the weak-scaling bench at its default shard (config 3 per GPU), then the human-scale geometry
# (config 4: 3.1 Gb genome, 26 + 43 GB filters replicated, 77.5 M reads per GPU)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r2_c6_gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c6_bench_n8.json 2> gpurun_out/r2_c6_bench_n8.err
echo "bench n8 rc=$?" >> gpurun_out/r2_c6_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --genome 387500000 --coverage 30 > gpurun_out/r2_c6_bench_human_n8.json 2> gpurun_out/r2_c6_bench_human_n8.err
echo "bench human rc=$?" >> gpurun_out/r2_c6_bench_human_n8.err
timeout 600 python -m pytest tests/test_cli_gpu.py -m gpu -q -k "sharded" > gpurun_out/r2_c6_pytest.log 2>&1
echo done
