#!/bin/bash
mkdir -p gpurun_out
NP=${1:-2}
( timeout 200 python -m pytest tests/test_gpu_filter.py -m gpu -q -k "label" ) > gpurun_out/c18_pytest.log 2>&1
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NP --steps 8 --warmup 3 --no-parity ) > gpurun_out/c18_bench_c4_np$NP.log 2>&1
tail -2 gpurun_out/c18_pytest.log
grep '^{"metric' gpurun_out/c18_bench_c4_np$NP.log | cut -c1-300
