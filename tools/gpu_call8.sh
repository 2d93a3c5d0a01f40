#!/bin/bash
# single GPU: new tests, pruning-knob probe, tracked bench lines for the other configs
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_filter.py tests/test_gpu_graph.py -m gpu -q ) > gpurun_out/c8_pytest.log 2>&1
( time timeout 600 python tools/r02_probe.py --config c4 --what prune ) > gpurun_out/c8_probe_prune.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/c8_bench_c4.log 2>&1
( time timeout 600 python bench.py --config c3 ) > gpurun_out/c8_bench_c3.log 2>&1
( time timeout 900 python bench.py --config c2 ) > gpurun_out/c8_bench_c2pca.log 2>&1
( time timeout 900 python bench.py --config c2 --n-pca none ) > gpurun_out/c8_bench_c2raw.log 2>&1
( time timeout 900 python bench.py --config c4iso --steps 3 --warmup 2 ) > gpurun_out/c8_bench_c4iso.log 2>&1
( time timeout 1200 python bench.py --config c5 --steps 3 --warmup 2 ) > gpurun_out/c8_bench_c5.log 2>&1
tail -2 gpurun_out/c8_pytest.log
grep "^{" gpurun_out/c8_probe_prune.log
for f in c4 c3 c2pca c2raw c4iso c5; do grep '^{"metric' gpurun_out/c8_bench_$f.log | cut -c1-260; done
