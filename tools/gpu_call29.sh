#!/bin/bash
# the round's last call: the whole GPU suite, smoke and the driver's bench command on the final tree
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2_c29_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c29_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c29_smoke.log 2>&1
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_c29_bench_c3.json 2> gpurun_out/r2_c29_bench_c3.err
echo "bench rc=$?" >> gpurun_out/r2_c29_bench_c3.err
echo done
