#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -q -x -k "golden or port or config3 or known or batch_split" ) > gpurun_out/r2_c13_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c13_pytest.log
VAR_GENOME=64000000 VAR_COVERAGE=30 VAR_ITERS=4 timeout 600 python tools/p3_variants.py main deep3 main:KBBQ_SCAN_DEEP=0 > gpurun_out/r2_c13_variants_c3.txt 2>&1
timeout 300 python tools/p3_variants.py main:KBBQ_SCAN_DEEP=1 deep3:KBBQ_SCAN_DEEP=1 main > gpurun_out/r2_c13_variants_c2.txt 2>&1
echo done
