#!/bin/bash
# one B200: bench lines of the other BASELINE configs and the adversarial workload
mkdir -p gpurun_out
( time timeout 900 python bench.py --config c5 --steps 3 --warmup 2 --no-cpu-baseline ) > gpurun_out/final_bench_c5.log 2>&1
( time timeout 600 python bench.py --config c3 --no-cpu-baseline ) > gpurun_out/final_bench_c3.log 2>&1
( time timeout 600 python bench.py --config c2 --no-cpu-baseline ) > gpurun_out/final_bench_c2pca.log 2>&1
for f in c5 c3 c2pca; do grep '^{"metric' gpurun_out/final_bench_$f.log | cut -c1-240; done
