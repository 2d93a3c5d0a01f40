#!/bin/bash
# GPU call 1 of round 2: tests, probe, bench, launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
nproc > gpurun_out/c1_nproc.txt
( time timeout 1500 python -m pytest tests -m gpu -q -k "not c5" --durations=15 ) > gpurun_out/c1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c1_pytest.log
( time MELD_B200_TIMING=1 timeout 600 python tools/r02_probe.py --config c4 ) > gpurun_out/c1_probe.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/c1_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/c1_launches.csv python bench.py --steps 1 --warmup 1 --no-parity --no-cpu-baseline > gpurun_out/c1_ncu_bench.log 2>&1
tail -5 gpurun_out/c1_pytest.log
tail -3 gpurun_out/c1_bench.log | cut -c1-600
