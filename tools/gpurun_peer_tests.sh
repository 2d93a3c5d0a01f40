#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_filter.py tests/test_gpu_graph.py -m gpu -q -k "peer or lanczos or dist or stage2" ) > gpurun_out/final_peer_tests.log 2>&1
tail -2 gpurun_out/final_peer_tests.log
