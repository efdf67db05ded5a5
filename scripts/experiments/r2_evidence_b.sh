#!/bin/bash
# round 2, second pass of this session: LayerNorm inside the residual GEMM launches + wide TMA boxes in the attention core
set -u
out=gpurun_out/r2b
mkdir -p $out
(timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "wide_boxes or inside_the_launch" 2>&1 | tail -12 | cut -c1-300)
(timeout 300 python -m pytest tests/test_model_gpu.py -q -m gpu -k "bitwise_neutral" 2>&1 | tail -8 | cut -c1-300)
(timeout 1500 python -m pytest tests -q -m gpu -x > $out/pytest_gpu.log 2>&1; tail -15 $out/pytest_gpu.log | cut -c1-300)
for w in 0 1 2; do
  echo "== kernels-only sdpa_wide=$w"; timeout 200 python bench.py --kernels-only --flag sdpa_wide=$w 2>&1 | grep -E "^(sdpa|gemm_out|gemm_ff2|layernorm|attn_block|wgrad)" | cut -c1-220
done
for w in 2 3; do
  echo "== kernels-only wgrad_pair=$w"; timeout 200 python bench.py --kernels-only --flag wgrad_pair=$w 2>&1 | grep -E "^wgrad"
done
for cfg in "sdpa_wide=0 gemm_fuse_ln=0" "sdpa_wide=0 gemm_fuse_ln=1" "sdpa_wide=1 gemm_fuse_ln=1" "sdpa_wide=2 gemm_fuse_ln=1"; do
  fl=""; for kv in $cfg; do fl="$fl --flag $kv"; done
  echo "== bench --no-extras $cfg"
  timeout 300 python bench.py --steps 10 --warmup 4 --no-extras $fl 2> $out/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['gpu_launches'], d['clocks'])" || tail -5 $out/err.txt
done
echo "== full bench (defaults)"
timeout 900 python bench.py --no-sweep > $out/bench.json 2> $out/bench.err; echo rc=$?
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2b/bench.json').read().strip().splitlines()[-1])
    print('AR', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['clocks'], d['gpu_launches'])
    print('b1', d.get('batch1')); print('kernels', d.get('kernels'))
    t=d.get('train') or {}; print('train', {k:t.get(k) for k in ['ms_per_step','samples_per_s','algorithmic_tflops_per_gpu','gpu_launches_per_step','roofline','clocks','error']})
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r2b/bench.err').read()[-2500:])
PY
