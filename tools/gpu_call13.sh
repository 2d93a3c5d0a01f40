#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_scale.py -m gpu -q -k "not c5 and not c2raw" ) > gpurun_out/c13_pytest.log 2>&1
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
) > gpurun_out/c13_timing.log 2>&1
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/c13_bench_c4.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c13_smoke.log 2>&1
tail -3 gpurun_out/c13_pytest.log
grep "refine" gpurun_out/c13_timing.log | tail -2
grep '^{"metric' gpurun_out/c13_bench_c4.log | cut -c1-260
tail -2 gpurun_out/c13_smoke.log
