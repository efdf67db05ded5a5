#!/bin/bash
# round 2, first GPU pass: new GPU tests, then the full bench line at N=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_model_gpu.py -x -q -m gpu -s 2>&1 | tail -40 > gpurun_out/r2_tests_a.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -5 gpurun_out/r2_tests_a.log
tail -c 3000 gpurun_out/r2_bench_n1.json
tail -5 gpurun_out/r2_bench_n1.err
