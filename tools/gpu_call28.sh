#!/bin/bash
set -x
mkdir -p gpurun_out
VAR_GENOME=64000000 VAR_COVERAGE=30 VAR_ITERS=4 timeout 600 python tools/p3_variants.py main prelong main prelong > gpurun_out/r2_c28_variants_c3.txt 2>&1
timeout 300 python tools/p3_variants.py main prelong > gpurun_out/r2_c28_variants_c2.txt 2>&1
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_cli_gpu.py -m gpu -q -x ) > gpurun_out/r2_c28_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c28_pytest.log
echo done
