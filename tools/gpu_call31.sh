#!/bin/bash
# long batches with their repair list split between the common and the long-read kernel
set -x
mkdir -p gpurun_out
( time timeout 240 python -m pytest tests/test_gpu_parity.py tests/test_cli_gpu.py -m gpu -q -x -k "long or beyond or longest or golden" ) > gpurun_out/r2_c31_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c31_pytest.log
echo done
