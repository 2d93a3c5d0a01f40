#!/bin/bash
mkdir -p gpurun_out
NP=${1:-2}
( timeout 300 python -m pytest tests/test_gpu_filter.py -m gpu -q -k "label or api or golden" ) > gpurun_out/c17_pytest.log 2>&1
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29520 tools/dist_check.py --config c4 --reps 3 ) > gpurun_out/c17_dist_check_np$NP.log 2>&1
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NP --steps 5 --warmup 3 ) > gpurun_out/c17_bench_c4_np$NP.log 2>&1
tail -3 gpurun_out/c17_pytest.log
grep "^{" gpurun_out/c17_dist_check_np$NP.log | tail -14
tail -5 gpurun_out/c17_dist_check_np$NP.log | grep -v "^{" | tail -5
grep '^{"metric' gpurun_out/c17_bench_c4_np$NP.log | cut -c1-330
tail -3 gpurun_out/c17_bench_c4_np$NP.log | grep -v '^{"metric' | cut -c1-300
