#!/bin/bash
set -x
mkdir -p gpurun_out
CLI_TIMING_EXIT_MODES=1,2,3,6,7 timeout 600 python tools/cli_timing.py 3 > gpurun_out/r2_c23_cli_exit.txt 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-file-leg --no-cpu-baseline > gpurun_out/r2_c23_bench_c3.json 2> gpurun_out/r2_c23_bench_c3.err
echo done
