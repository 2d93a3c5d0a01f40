#!/bin/bash
# GPU call 3 (2 GPUs): multi-GPU correctness + timing
mkdir -p gpurun_out
NP=${1:-2}
( timeout 300 python -m pytest tests/test_gpu_filter.py tests/test_gpu_graph.py -m gpu -q -k "kat or lmax or decay or dist or peer or lanczos or two_stage" ) > gpurun_out/c3_pytest.log 2>&1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py --config c4 ) > gpurun_out/c3_dist_check_np$NP.log 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NP --steps 5 --warmup 3 ) > gpurun_out/c3_bench_np$NP.log 2>&1
( time MELD_B200_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29513 tools/dist_check.py --config c4 --reps 1 ) > gpurun_out/c3_dist_timing_np$NP.log 2>&1
tail -3 gpurun_out/c3_pytest.log
grep -v "^\[" gpurun_out/c3_dist_check_np$NP.log | tail -20
grep '^{"metric' gpurun_out/c3_bench_np$NP.log | cut -c1-300
