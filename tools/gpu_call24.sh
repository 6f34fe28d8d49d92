#!/bin/bash
set -x
mkdir -p gpurun_out
VAR_GENOME=64000000 VAR_COVERAGE=30 VAR_ITERS=4 timeout 900 python tools/p3_variants.py main p3coop4 p3coop2 > gpurun_out/r2_c24_variants_c3.txt 2>&1
timeout 300 python tools/p3_variants.py main p3coop4 > gpurun_out/r2_c24_variants_c2.txt 2>&1
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -q -x ) > gpurun_out/r2_c24_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c24_pytest.log
echo done
