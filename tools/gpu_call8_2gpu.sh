#!/bin/bash
# 2 GPUs: the NVSwitch-multicast merge (NVLS) -- parity first, then a short bench with it and with the peer-memory merge
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_mgpu.py -m gpu -q ) > gpurun_out/r2_c8_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c8_pytest.log
KBBQ_NVLS_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_c8_bench_n2_nvls.json 2> gpurun_out/r2_c8_bench_n2_nvls.err
echo "bench nvls rc=$?" >> gpurun_out/r2_c8_bench_n2_nvls.err
KBBQ_MERGE=p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c8_bench_n2_p2p.json 2> gpurun_out/r2_c8_bench_n2_p2p.err
echo "bench p2p rc=$?" >> gpurun_out/r2_c8_bench_n2_p2p.err
echo done
