#!/bin/bash
# Last call of the round: the whole GPU suite and smoke() on the committed code.
mkdir -p gpurun_out
( timeout 230 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02j_gpu_tests.log 2>&1
tail -2 gpurun_out/r02j_gpu_tests.log
( timeout 40 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/r02j_smoke.log 2>&1
tail -1 gpurun_out/r02j_smoke.log
