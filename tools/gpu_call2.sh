#!/bin/bash
# round-2 GPU call 2: parity suite (incl. config-3 digests), benches, per-kernel ncu
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_c2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c2_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c2_bench_c3.json 2> gpurun_out/r2_c2_bench_c3.err
echo "bench rc=$?" >> gpurun_out/r2_c2_bench_c3.err
timeout 300 python bench.py --steps 10 --warmup 3 --genome 4600000 --coverage 50 --no-cpu-baseline > gpurun_out/r2_c2_bench_c2.json 2> gpurun_out/r2_c2_bench_c2.err
timeout 300 python tools/one_step.py 2 64000000 30 > gpurun_out/r2_c2_one_step.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum \
    --clock-control none --csv -k regex:'covariate|pass4|pass3|pass1' --log-file gpurun_out/r2_c2_launches_c3.csv python tools/one_step.py 2 64000000 30 > gpurun_out/r2_c2_ncu.log 2>&1
echo done
