#!/bin/bash
# Round-end evidence, run on the GPU box AFTER the plain runs have exited 0 (numbers printed under ncu are never bench
# values):   tools/profile_round.sh TAG
#   gpurun_out/TAG_bench.json       plain bench line
#   gpurun_out/TAG_launches.csv     ncu launch list of `bench.py --steps 2 --warmup 3` (gpu__time_duration.sum per launch)
#   gpurun_out/TAG_full.ncu-rep     ncu --set full of one whole step (6 kernels), source imported
set -e
TAG=${1:-r1}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
python tools/one_step.py 3 > gpurun_out/${TAG}_one_step.log 2>&1
ncu --set full --import-source on --clock-control none --launch-skip 9 --launch-count 6 -f -o gpurun_out/${TAG}_full \
    python tools/one_step.py 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_one_step.log
