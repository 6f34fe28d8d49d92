#!/bin/bash
# final validation on 2 GPUs: the whole -m gpu suite (multi-GPU tests included), a 2-GPU bench, the 1-GPU bench line
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2_c11_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c11_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c11_bench_n2.json 2> gpurun_out/r2_c11_bench_n2.err
echo "bench n2 rc=$?" >> gpurun_out/r2_c11_bench_n2.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-file-leg > gpurun_out/r2_c11_bench_c3.json 2> gpurun_out/r2_c11_bench_c3.err
echo done
