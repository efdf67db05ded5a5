#!/bin/bash
# last GPU seconds of the round: launch lists of one B=128 frame-step and one training step with the final kernels
out=gpurun_out/r2z
mkdir -p $out
timeout 55 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras > $out/launches_bench.log 2>&1
python scripts/summarize_launches.py $out/launches_bench.csv --between step_inc_kernel > $out/launches_bench_summary.txt 2>&1; head -12 $out/launches_bench_summary.txt
timeout 55 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/launches_train.csv python scripts/bench_train.py --steps 1 --warmup 1 > $out/launches_train.log 2>&1
python scripts/summarize_launches.py $out/launches_train.csv --between adam_kernel > $out/launches_train_summary.txt 2>&1; head -22 $out/launches_train_summary.txt
