#!/bin/bash
# round 2, 5th GPU pass: wgrad orientation tests + training tests, then the bench line (train leg with the new wgrad)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_train_gpu.py -q -m gpu -x -s 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2_tests_wgrad.log
tail -8 gpurun_out/r2_tests_wgrad.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-sweep --no-cpu > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1_b.json').read().strip().splitlines()[-1])
print('AR', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'b1', d.get('batch1'))
t=d['train']; print('train', {k:t.get(k) for k in ['ms_per_step','samples_per_s','algorithmic_tflops_per_gpu','gpu_launches_per_step','roofline','clocks','error']})
PY
tail -3 gpurun_out/r2_bench_n1_b.err
