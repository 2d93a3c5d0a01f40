#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/c12_pytest.log 2>&1
( time MELD_B200_TIMING=1 timeout 300 python - <<'PY'
import sys
sys.path.insert(0, '.')
import torch, meld_b200
from meld_b200 import synthetic
X, y, kw = synthetic.make_config("c4")
Xd = torch.from_numpy(X).cuda()
for rep in range(3):
    sys.stderr.write("=== build %d\n" % rep); sys.stderr.flush()
    g = meld_b200.DeviceGraph.from_data(Xd, knn=15)
    torch.cuda.synchronize()
PY
) > gpurun_out/c12_timing.log 2>&1
( time timeout 600 python tools/probe_e2e_outliers.py ) > gpurun_out/c12_e2e_outliers.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/c12_bench_c4.log 2>&1
tail -3 gpurun_out/c12_pytest.log
grep "refine\|cell order\|candidate search total" gpurun_out/c12_timing.log | tail -4
grep "^{" gpurun_out/c12_e2e_outliers.log | cut -c1-420
grep '^{"metric' gpurun_out/c12_bench_c4.log | cut -c1-260
