#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_c4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c4_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c4_bench_c3.json 2> gpurun_out/r2_c4_bench_c3.err
echo "bench rc=$?" >> gpurun_out/r2_c4_bench_c3.err
timeout 300 python bench.py --steps 10 --warmup 3 --genome 4600000 --coverage 50 --no-cpu-baseline > gpurun_out/r2_c4_bench_c2.json 2> gpurun_out/r2_c4_bench_c2.err
VAR_GENOME=64000000 VAR_COVERAGE=30 VAR_ITERS=4 timeout 600 python tools/p3_variants.py main main:KBBQ_NO_PREFETCH=1 main:KBBQ_INSERT_MODE=1 > gpurun_out/r2_c4_variants_c3.txt 2>&1
timeout 300 python tools/one_step.py 2 64000000 30 > gpurun_out/r2_c4_one_step.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum \
    --clock-control none --csv --log-file gpurun_out/r2_c4_launches_c3.csv python tools/one_step.py 2 64000000 30 > gpurun_out/r2_c4_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pass3_repair -s 1 -c 1 -o gpurun_out/r2_c4_repair_c2 python tools/one_step.py 2 > gpurun_out/r2_c4_ncu_full.log 2>&1
echo done
