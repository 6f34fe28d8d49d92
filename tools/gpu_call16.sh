#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_cli_gpu.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r2_c16_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c16_pytest.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_c16_bench_c3.json 2> gpurun_out/r2_c16_bench_c3.err
echo "bench rc=$?" >> gpurun_out/r2_c16_bench_c3.err
timeout 300 python tools/one_step.py 2 64000000 30 > gpurun_out/r2_c16_one_step_c3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pass3_repair -s 1 -c 1 -o gpurun_out/r2_c16_repair_c3 python tools/one_step.py 2 64000000 30 > gpurun_out/r2_c16_ncu.log 2>&1
echo done
