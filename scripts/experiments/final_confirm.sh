#!/bin/bash
set -u
out=gpurun_out/final3
mkdir -p $out
(timeout 240 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -q -m gpu -x > $out/pytest.log 2>&1; tail -2 $out/pytest.log)
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 120 python scripts/sweep_batch.py --modes precise --batches 1,2,4,128 --steps 12 > $out/sweep_small.json 2>/dev/null; grep frames_per_s $out/sweep_small.json | cut -c1-110
