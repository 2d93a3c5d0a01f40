#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -k "not c5 and not c2raw" --durations=8 ) > gpurun_out/c11_pytest.log 2>&1
( time PROBE_PRUNE_SHORT=1 timeout 600 python tools/r02_probe.py --config c4 --what prune ) > gpurun_out/c11_probe_prune.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/c11_bench_c4.log 2>&1
( time timeout 900 python bench.py --config c2 --no-cpu-baseline ) > gpurun_out/c11_bench_c2pca.log 2>&1
( time timeout 1200 python bench.py --config c5 --steps 3 --warmup 2 --no-cpu-baseline ) > gpurun_out/c11_bench_c5.log 2>&1
tail -3 gpurun_out/c11_pytest.log
grep "^{" gpurun_out/c11_probe_prune.log
for f in c4 c2pca c5; do grep '^{"metric' gpurun_out/c11_bench_$f.log | cut -c1-260; done
