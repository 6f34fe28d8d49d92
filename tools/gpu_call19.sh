#!/bin/bash
set -x
mkdir -p gpurun_out
VAR_GENOME=64000000 VAR_COVERAGE=30 VAR_ITERS=4 timeout 900 python tools/p3_variants.py main p2b3 p2b3:KBBQ_P2_DEEP=1 > gpurun_out/r2_c19_variants_c3.txt 2>&1
echo done
