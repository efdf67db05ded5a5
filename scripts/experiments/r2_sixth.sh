#!/bin/bash
# round 2, 6th GPU pass: pipelined attention backward -- tests, then A/B timing at the training shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_gpu.py -q -m gpu -x -k "sdpa" 2>&1 | tail -12 > gpurun_out/r2_tests_sdpabwd.log
tail -5 gpurun_out/r2_tests_sdpabwd.log
timeout 300 python scripts/experiments/time_sdpa_bwd.py 2>&1 | tail -8
