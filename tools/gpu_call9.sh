#!/bin/bash
mkdir -p gpurun_out
NP=${1:-2}
( timeout 300 python tools/bench_sweep.py ) > gpurun_out/c15_sweep.log 2>&1
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29520 tools/dist_check.py --config c4 --reps 3 ) > gpurun_out/c15_dist_check_np$NP.log 2>&1
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $NP --steps 5 --warmup 3 ) > gpurun_out/c15_bench_c4_np$NP.log 2>&1
tail -3 gpurun_out/c15_pytest.log
grep "^{" gpurun_out/c15_dist_check_np$NP.log | tail -14
tail -5 gpurun_out/c15_dist_check_np$NP.log | grep -v "^{" | tail -5
grep '^{"metric' gpurun_out/c15_bench_c4_np$NP.log | cut -c1-330
tail -3 gpurun_out/c15_bench_c4_np$NP.log | grep -v '^{"metric' | cut -c1-300
