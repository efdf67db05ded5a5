#!/bin/bash
# last GPU call of the round: confirm the shipped defaults and refresh the headline lines
set -u
out=gpurun_out/final2
mkdir -p $out
(timeout 600 python -m pytest tests -q -m gpu > $out/pytest_gpu.log 2>&1; tail -2 $out/pytest_gpu.log)
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > $out/bench.json 2> $out/bench.err; tail -1 $out/bench.json | cut -c1-330
timeout 200 python scripts/bench_train.py --steps 6 --warmup 3 > $out/train.json 2>/dev/null; tail -1 $out/train.json | cut -c1-300
timeout 300 python scripts/sweep_batch.py > $out/sweep.json 2> $out/sweep.err; grep -c frames_per_s $out/sweep.json
FACT_FLAGS=gemm_finish_ln=1 timeout 200 python scripts/sweep_batch.py --modes precise --batches 1,2,4 --steps 24 > $out/sweep_finish_ln.json 2>/dev/null
grep frames_per_s $out/sweep_finish_ln.json | cut -c1-120
