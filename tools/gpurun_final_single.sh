#!/bin/bash
# final single-GPU check: what the driver runs at round end
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/final_pytest.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final_smoke.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/final_bench_c4.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/final_ref.log 2>&1
tail -3 gpurun_out/final_pytest.log
tail -1 gpurun_out/final_smoke.log | cut -c1-200
grep '^{"metric' gpurun_out/final_bench_c4.log | cut -c1-260
grep '^{"impl' gpurun_out/final_ref.log | cut -c1-260
