#!/bin/bash
# validation of the final tree (1 GPU): parity suite, the bench line as the driver asks for it, smoke
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2_c10_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c10_pytest.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_c10_bench_c3.json 2> gpurun_out/r2_c10_bench_c3.err
echo "bench rc=$?" >> gpurun_out/r2_c10_bench_c3.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c10_smoke.log 2>&1
echo done
