#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py tests/test_synth.py -m gpu -q ) > gpurun_out/r2_c12_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c12_pytest.log
VAR_GENOME=64000000 VAR_COVERAGE=30 VAR_ITERS=4 timeout 600 python tools/p3_variants.py main occ5 occ3 > gpurun_out/r2_c12_variants_c3.txt 2>&1
timeout 300 python tools/p3_variants.py main occ5 occ3 > gpurun_out/r2_c12_variants_c2.txt 2>&1
timeout 300 python tools/one_step.py 1 64000000 30 > gpurun_out/r2_c12_one_step.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pass3_scan -c 1 -o gpurun_out/r2_c12_scan_c3 python tools/one_step.py 1 64000000 30 > gpurun_out/r2_c12_ncu_full.log 2>&1
echo done
