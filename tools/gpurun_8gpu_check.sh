#!/bin/bash
# 8-GPU box: the strong-scaling bench line of config 4 (what the driver's scaling run does at N = 8)
mkdir -p gpurun_out
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/final_bench_c4_np8.log 2>&1
grep '^{"metric' gpurun_out/final_bench_c4_np8.log | cut -c1-300
tail -4 gpurun_out/final_bench_c4_np8.log | grep -v '^{"metric' | cut -c1-200
