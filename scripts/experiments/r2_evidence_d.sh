#!/bin/bash
# round 2, fourth pass: softmax that reads half of the scores once, LayerNorm warps dealt over all clusters of a row block
set -u
out=gpurun_out/r2d
mkdir -p $out
(timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "issue_order or inside_the_launch or test_sdpa" 2>&1 | tail -12 | cut -c1-300)
(timeout 300 python -m pytest tests/test_model_gpu.py -q -m gpu -k "bitwise_neutral or parity or reference" 2>&1 | tail -8 | cut -c1-300)
for cfg in "sdpa_pipe=0" "sdpa_pipe=1"; do
  echo "== kernels-only $cfg"; FACT_FLAGS=$cfg timeout 200 python bench.py --kernels-only 2>&1 | grep -E "^(sdpa|gemm_out|gemm_ff2|layernorm|attn_block)" | cut -c1-200
done
echo "== kernels-only bf16 mode"; timeout 200 python bench.py --kernels-only --mode bf16 2>&1 | grep -E "^(sdpa|gemm_out|gemm_ff2|layernorm)" | cut -c1-200
for cfg in "gemm_fuse_ln=0" "gemm_fuse_ln=1" "gemm_fuse_ln=0" "gemm_fuse_ln=1"; do
  echo "== bench --no-extras $cfg"
  FACT_FLAGS=$cfg timeout 300 python bench.py --steps 10 --warmup 4 --no-extras 2> $out/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['gpu_launches'], d['clocks'])" || tail -5 $out/err.txt
done
for cfg in "gemm_fuse_ln=0" "gemm_fuse_ln=1"; do
  echo "== train $cfg"; FACT_FLAGS=$cfg timeout 300 python scripts/bench_train.py --steps 8 --warmup 3 2>/dev/null | tail -1 | cut -c1-330
done
(timeout 1500 python -m pytest tests -q -m gpu -x > $out/pytest_gpu.log 2>&1; tail -4 $out/pytest_gpu.log | cut -c1-300)
