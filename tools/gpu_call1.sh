#!/bin/bash
# round-2 GPU call 1: parity suite, lab micro-benchmarks, config-3 bench, per-kernel ncu (time + DRAM bytes)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c1_pytest.log
timeout 300 ./build/probe_lab > gpurun_out/r2_probe_lab.json 2> gpurun_out/r2_probe_lab.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_c1_bench_c3.json 2> gpurun_out/r2_c1_bench_c3.err
echo "bench rc=$?" >> gpurun_out/r2_c1_bench_c3.err
timeout 300 python bench.py --steps 10 --warmup 3 --genome 4600000 --coverage 50 --no-cpu-baseline > gpurun_out/r2_c1_bench_c2.json 2> gpurun_out/r2_c1_bench_c2.err
VAR_GENOME=64000000 VAR_COVERAGE=30 VAR_ITERS=4 timeout 600 python tools/p3_variants.py p3r3 p3r5 > gpurun_out/r2_c1_variants_c3.txt 2>&1
timeout 300 python tools/p3_variants.py p3r3 p3r5 > gpurun_out/r2_c1_variants_c2.txt 2>&1
# ncu: lab (DRAM bytes per probe variant), then one config-3 step (device time + DRAM bytes of every launch)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --csv \
    --log-file gpurun_out/r2_probe_lab_ncu.csv ./build/probe_lab quick > gpurun_out/r2_probe_lab_ncu.log 2>&1
timeout 300 python tools/one_step.py 2 64000000 30 > gpurun_out/r2_c1_one_step.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum \
    --clock-control none --csv --log-file gpurun_out/r2_c1_launches_c3.csv python tools/one_step.py 2 64000000 30 > gpurun_out/r2_c1_ncu.log 2>&1
echo done
