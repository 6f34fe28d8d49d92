#!/bin/bash
# final round-2 measurement call (1 GPU): the whole GPU suite, the driver's bench commands, ncu launch lists, smoke
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2_c20_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c20_pytest.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_c20_bench_c3.json 2> gpurun_out/r2_c20_bench_c3.err
echo "bench rc=$?" >> gpurun_out/r2_c20_bench_c3.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_c20_bench_ref.json 2> gpurun_out/r2_c20_bench_ref.err
timeout 300 python bench.py --steps 20 --warmup 5 --genome 4600000 --coverage 50 --no-cpu-baseline --no-file-leg > gpurun_out/r2_c20_bench_c2.json 2> gpurun_out/r2_c20_bench_c2.err
timeout 300 python tools/one_step.py 2 64000000 30 > gpurun_out/r2_c20_one_step_c3.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum \
    --clock-control none --csv --log-file gpurun_out/r2_c20_launches_c3.csv python tools/one_step.py 2 64000000 30 > gpurun_out/r2_c20_ncu_c3.log 2>&1
timeout 300 python tools/one_step.py 3 > gpurun_out/r2_c20_one_step_c2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum \
    --clock-control none --csv --log-file gpurun_out/r2_c20_launches_c2.csv python tools/one_step.py 3 > gpurun_out/r2_c20_ncu_c2.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c20_smoke.log 2>&1
echo done
