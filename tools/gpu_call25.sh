#!/bin/bash
# the driver's commands on the final tree: both arms at N=1, then the ncu launch list of the bench command itself
set -x
mkdir -p gpurun_out
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_c25_bench_ref.json 2> gpurun_out/r2_c25_bench_ref.err
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_c25_bench_c3.json 2> gpurun_out/r2_c25_bench_c3.err
echo "bench rc=$?" >> gpurun_out/r2_c25_bench_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_c25_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-file-leg --no-e2e > gpurun_out/r2_c25_ncu_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c25_smoke.log 2>&1
echo done
