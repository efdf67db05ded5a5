#!/bin/bash
set -u
out=gpurun_out/exp_quick
mkdir -p $out
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_train_gpu.py tests/test_kernels_gpu.py -x -q -m gpu > $out/pytest.log 2>&1
tail -3 $out/pytest.log
timeout 300 python scripts/bench_train.py --steps 4 --warmup 2 2>/dev/null | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches_train.csv python scripts/bench_train.py --steps 1 --warmup 1 > $out/train_ncu.log 2>&1
python scripts/summarize_launches.py $out/launches_train.csv --between adam_kernel 2>/dev/null | head -16
