#!/bin/bash
# reads beyond 512 bases: the new long-read tests first, then the timing of the common path, then the rest of the parity suite
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "long or beyond or longest" ) > gpurun_out/r2_c26_pytest_long.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c26_pytest_long.log
VAR_GENOME=64000000 VAR_COVERAGE=30 VAR_ITERS=4 timeout 600 python tools/p3_variants.py main > gpurun_out/r2_c26_variants_c3.txt 2>&1
timeout 300 python tools/p3_variants.py main > gpurun_out/r2_c26_variants_c2.txt 2>&1
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -q -k "not (long or beyond or longest)" ) > gpurun_out/r2_c26_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c26_pytest.log
echo done
