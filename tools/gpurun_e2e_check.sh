#!/bin/bash
# After the DataFrame / pinned-result change: the filter + API tests and one default bench line.
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_gpu_filter.py -m gpu -x -q ) > gpurun_out/r02j_gpu_filter_tests.log 2>&1
tail -2 gpurun_out/r02j_gpu_filter_tests.log
( timeout 120 python bench.py ) > gpurun_out/r02j_bench_c4.jsonl 2> gpurun_out/r02j_bench_c4.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02j_bench_c4.jsonl").read().strip().splitlines()[-1])
print(l["ms_per_step"], l["e2e"]["step_ms"], l["e2e"].get("pageable_input_step_ms"), l["e2e"]["host_timings_ms_last_step"], l["parity"]["full_size_filter"]["normwise"])
PY
