#!/bin/bash
# round 2, sixth pass: attention output through shared memory + bulk tensor stores
set -u
out=gpurun_out/r2f
mkdir -p $out
(timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_backward_gpu.py -q -m gpu -k "sdpa" 2>&1 | tail -8 | cut -c1-300)
(timeout 300 python -m pytest tests/test_model_gpu.py -q -m gpu 2>&1 | tail -5 | cut -c1-300)
echo "== kernels-only precise"; timeout 200 python bench.py --kernels-only 2>&1 | grep -E "^(sdpa|attn_block)" | cut -c1-200
echo "== kernels-only bf16"; timeout 200 python bench.py --kernels-only --mode bf16 2>&1 | grep -E "^(sdpa)" | cut -c1-200
for i in 1 2; do
  echo "== bench --no-extras"
  timeout 300 python bench.py --steps 10 --warmup 4 --no-extras 2> $out/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['gpu_launches'], d['clocks'])" || tail -5 $out/err.txt
done
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:sdpa_tc_kernel -s 5 -c 1 -o $out/sdpa python bench.py --kernels-only > $out/sdpa.log 2>&1; tail -1 $out/sdpa.log
(timeout 1500 python -m pytest tests -q -m gpu -x > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log | cut -c1-300)
