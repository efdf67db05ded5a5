#!/bin/bash
# round 2, 7th GPU pass: full GPU suite, bench line at N=1 (train leg with all training-kernel changes)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r2_tests_full2.log
tail -3 gpurun_out/r2_tests_full2.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-sweep --no-cpu > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1_c.json').read().strip().splitlines()[-1])
print('AR', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'b1', d.get('batch1',{}).get('value'))
t=d['train']; print('train', {k:t.get(k) for k in ['ms_per_step','samples_per_s','algorithmic_tflops_per_gpu','gpu_launches_per_step','clocks','error']})
PY
tail -3 gpurun_out/r2_bench_n1_c.err
