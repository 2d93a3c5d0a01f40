#!/bin/bash
# 8-GPU box: validate the peer-store path at 8 and 4 ranks, strong-scaling bench lines
mkdir -p gpurun_out
for NP in 8 4; do
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 2951$NP bench.py --gpus $NP --steps 5 --warmup 3 --no-parity ) > gpurun_out/c5_bench_np$NP.log 2>&1
done
( time MELD_B200_TIMING=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 tools/dist_check.py --config c4 --reps 3 ) > gpurun_out/c5_dist_check_np8.log 2>&1
grep -v "^\[" gpurun_out/c5_dist_check_np8.log | grep "^{" | tail -12
for NP in 8 4; do grep '^{"metric' gpurun_out/c5_bench_np$NP.log | cut -c1-330; done
