#!/bin/bash
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c32_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_c32_smoke.log
( timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or 700" ) > gpurun_out/r2_c32_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c32_pytest.log
echo done
