#!/bin/bash
# Launch list of the final code: one device-arm step + one warm-up, no CPU legs.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-parity --no-cpu-baseline"
( timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02i_launches_bench_c4.csv $CMD ) > gpurun_out/r02i_launches.log 2>&1
tail -c 400 gpurun_out/r02i_launches.log
wc -l gpurun_out/r02i_launches_bench_c4.csv
