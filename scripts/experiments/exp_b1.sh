#!/bin/bash
set -u
out=gpurun_out/exp_b1
mkdir -p $out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q -m gpu > $out/pytest.log 2>&1
tail -3 $out/pytest.log
for f in 0 1; do
  echo "== gemm_finish_ln=$f"
  FACT_FLAGS=gemm_finish_ln=$f timeout 300 python scripts/sweep_batch.py --modes precise --batches 1,2,4,8 --steps 24 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        r=json.loads(l); print(r['batch'], round(r['frames_per_s'],1), round(r['ms_per_frame_step'],3))"
done
timeout 300 python bench.py --steps 8 --warmup 3 --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; r=json.loads(sys.stdin.read()); print('bench', r['value'], r['ms_per_step'], r.get('e2e',{}).get('value'), r.get('gpu_launches'))"
