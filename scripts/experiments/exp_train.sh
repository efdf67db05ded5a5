#!/bin/bash
# training-path check: unit tests of the backward blocks + train-step timing (+ optional ncu launch list)
set -u
out=gpurun_out/exp_train
mkdir -p $out
timeout 1200 python -m pytest tests -x -q -m gpu > $out/pytest.log 2>&1
tail -5 $out/pytest.log
timeout 300 python scripts/bench_train.py --steps 4 --warmup 2 2>/dev/null | tail -1
if [ "${1:-}" = "ncu" ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $out/launches_train.csv python scripts/bench_train.py --steps 1 --warmup 1 > $out/train_ncu.log 2>&1
  tail -1 $out/train_ncu.log
fi
