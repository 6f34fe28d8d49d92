#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/cli_timing.py 3 > gpurun_out/r2_c18_cli_timing.txt 2>&1
( time timeout 600 python -m pytest tests/test_cli_gpu.py -m gpu -q -x ) > gpurun_out/r2_c18_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c18_pytest.log
echo done
