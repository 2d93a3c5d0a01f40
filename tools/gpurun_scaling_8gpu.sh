#!/bin/bash
mkdir -p gpurun_out
for NP in 8 4; do
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 2953$NP bench.py --gpus $NP --steps 5 --warmup 3 ) > gpurun_out/c16_bench_c4_np$NP.log 2>&1
done
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --steps 3 --warmup 2 --config c5 --no-parity ) > gpurun_out/c16_bench_c5_np8.log 2>&1
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --impl reference --steps 2 --warmup 1 ) > gpurun_out/c16_ref_np8.log 2>&1
for NP in 8 4; do grep '^{"metric' gpurun_out/c16_bench_c4_np$NP.log | cut -c1-300; done
grep '^{"metric' gpurun_out/c16_bench_c5_np8.log | cut -c1-300
grep '^{"impl' gpurun_out/c16_ref_np8.log | cut -c1-300
