#!/bin/bash
set -u
out=gpurun_out/exp_sdpa_fwd
mkdir -p $out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_backward_gpu.py -x -q -m gpu -k "sdpa" > $out/pytest.log 2>&1
tail -4 $out/pytest.log
if grep -q "failed\|error" $out/pytest.log; then grep -n "Error\|assert" $out/pytest.log | head -10; exit 1; fi
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_train_gpu.py -x -q -m gpu > $out/pytest_model.log 2>&1
tail -3 $out/pytest_model.log
timeout 300 python bench.py --kernels-only 2>/dev/null | grep "sdpa\|attn"
timeout 300 python bench.py --kernels-only --mode bf16 2>/dev/null | grep "sdpa"
timeout 300 python bench.py --steps 8 --warmup 3 --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; r=json.loads(sys.stdin.read()); print('bench', r['value'], r['ms_per_step'], r.get('e2e',{}).get('value'), r.get('clocks'))"
timeout 300 python scripts/bench_train.py --steps 4 --warmup 2 2>/dev/null | tail -1 | cut -c1-260
