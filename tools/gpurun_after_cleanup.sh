#!/bin/bash
# After removing the staged Chebyshev kernel: the whole GPU suite, smoke() and one default bench line.
mkdir -p gpurun_out
( timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02i_gpu_tests.log 2>&1
tail -3 gpurun_out/r02i_gpu_tests.log
( timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/r02i_smoke.log 2>&1
tail -1 gpurun_out/r02i_smoke.log
( timeout 200 python bench.py ) > gpurun_out/r02i_bench_c4.jsonl 2> gpurun_out/r02i_bench_c4.err
tail -c 1500 gpurun_out/r02i_bench_c4.jsonl
