#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_filter.py -m gpu -q -k "peer or lanczos" ) > gpurun_out/c14_pytest.log 2>&1
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 tools/dist_check.py --config c4 --reps 3 ) > gpurun_out/c14_dist_check_np8.log 2>&1
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-parity ) > gpurun_out/c14_bench_c4_np8.log 2>&1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 2 --no-parity --config c5 ) > gpurun_out/c14_bench_c5_np8.log 2>&1
tail -2 gpurun_out/c14_pytest.log
grep "^{" gpurun_out/c14_dist_check_np8.log | tail -9
grep '^{"metric' gpurun_out/c14_bench_c4_np8.log | cut -c1-330
grep '^{"metric' gpurun_out/c14_bench_c5_np8.log | cut -c1-330
