#!/bin/bash
# final tree on 2 GPUs: the multi-GPU tests (torchrun, all merge transports), the sharded CLI test, a 2-GPU bench line
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_mgpu.py tests/test_cli_gpu.py tests/test_gpu_parity.py -m gpu -q ) > gpurun_out/r2_c21_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c21_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c21_bench_n2.json 2> gpurun_out/r2_c21_bench_n2.err
echo "bench n2 rc=$?" >> gpurun_out/r2_c21_bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-file-leg > gpurun_out/r2_c21_bench_c3.json 2> gpurun_out/r2_c21_bench_c3.err
echo done
