#!/bin/bash
# round 2, third GPU pass: whole GPU suite, then a launch list of one batch-1 frame
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_tests_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_b1_launches.csv python scripts/sweep_batch.py --modes precise --batches 1 --steps 2 --warmup 1 > gpurun_out/r2_b1_ncu.log 2>&1
tail -5 gpurun_out/r2_tests_full.log
python scripts/summarize_launches.py gpurun_out/r2_b1_launches.csv --between step_inc_kernel
