#!/bin/bash
# round 2: the bench line under torchrun at N = 2 with this session's kernels (AR replicas, sweep, data-parallel training leg)
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --sweep 1,16,128,512 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo rc=$?
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
    print('AR', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['clocks'])
    t=d['train']; print('train', {k:t.get(k) for k in ['ms_per_step','samples_per_s','allreduce_ms','exposed_allreduce_ms','ms_per_step_without_allreduce','ms_per_step_free_running','sync_skew_ms','allreduce_overlap','fused_step','gpu_topology','error']})
    print('modes', t.get('modes'))
    print([(p['batch_per_gpu'], round(p['frames_per_s'])) for p in d['batch_sweep']['points']])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r2_bench_n2.err').read()[-3000:])
PY
tail -3 gpurun_out/r2_bench_n2.err | cut -c1-300
timeout 300 python -m pytest tests/test_dp_fused_gpu.py tests/test_trainer_gloo.py -q 2>&1 | tail -3 | cut -c1-200
