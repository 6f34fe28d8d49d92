#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2_c3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c3_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c3_bench_c3.json 2> gpurun_out/r2_c3_bench_c3.err
echo "bench rc=$?" >> gpurun_out/r2_c3_bench_c3.err
timeout 300 python bench.py --steps 10 --warmup 3 --genome 4600000 --coverage 50 --no-cpu-baseline > gpurun_out/r2_c3_bench_c2.json 2> gpurun_out/r2_c3_bench_c2.err
echo done
VAR_GENOME=64000000 VAR_COVERAGE=30 VAR_ITERS=4 timeout 600 python tools/p3_variants.py main p3coop8 main:KBBQ_NO_KNOWN_BITS=1 > gpurun_out/r2_c3_variants_c3.txt 2>&1
timeout 300 python tools/p3_variants.py main p3coop8 main:KBBQ_NO_KNOWN_BITS=1 > gpurun_out/r2_c3_variants_c2.txt 2>&1
echo done2
