#!/bin/bash
# GPU call 2 of round 2 (1 GPU): full tests incl. c5, stage timing, lanczos history, bench, ncu captures.
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --durations=12 ) > gpurun_out/c2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c2_pytest.log
( MELD_B200_TIMING=1 MELD_B200_LANCZOS_DEBUG=1 timeout 300 python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import torch, meld_b200
from meld_b200 import synthetic
X, y, kw = synthetic.make_config("c4")
Xd = torch.from_numpy(X).cuda()
for rep in range(3):
    print("=== build", rep, flush=True); sys.stderr.write("=== build %d\n" % rep); sys.stderr.flush()
    g = meld_b200.DeviceGraph.from_data(Xd, knn=15)
    torch.cuda.synchronize()
print("=== lmax", flush=True); sys.stderr.write("=== lmax\n")
for tol in (1e-3, 1e-5):
    g._lmax = None
    t = time.perf_counter(); l = g.estimate_lmax(rel_tol=tol); dt = time.perf_counter() - t
    print("tol", tol, "lmax", l, "iters", g.lmax_iters, "ms", 1e3 * dt, flush=True)
PY
) > gpurun_out/c2_timing.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/c2_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cheby_flat2 -s 70 -c 2 -o gpurun_out/c2_flat2 python bench.py --steps 1 --warmup 1 --no-parity --no-cpu-baseline > gpurun_out/c2_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_search -s 3 -c 3 -o gpurun_out/c2_tcsearch python bench.py --steps 1 --warmup 1 --no-parity --no-cpu-baseline > gpurun_out/c2_ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/c2_launches.csv python bench.py --steps 1 --warmup 1 --no-parity --no-cpu-baseline > gpurun_out/c2_ncu3.log 2>&1
tail -4 gpurun_out/c2_pytest.log
grep '^{"metric' gpurun_out/c2_bench.log | cut -c1-400
