#!/bin/bash
# round 2: fused data-parallel step per slice under the backward, 2 GPUs
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_dp_fused_gpu.py -q -m gpu -s 2>&1 | tail -6 | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 12 --warmup 3 --no-sweep > gpurun_out/r2_bench_n2_staged.json 2> gpurun_out/r2_bench_n2_staged.err
echo rc=$?
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n2_staged.json').read().strip().splitlines()[-1])
    print('AR', round(d['value']), d['ms_per_step'])
    t=d['train']; print('train', {k:t.get(k) for k in ['ms_per_step','samples_per_s','exposed_allreduce_ms','ms_per_step_without_allreduce','ms_per_step_free_running','sync_skew_ms','allreduce_overlap','fused_step','error']})
    for k,v in (t.get('modes') or {}).items(): print(k, {a:round(b,3) for a,b in v.items()})
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r2_bench_n2_staged.err').read()[-3000:])
PY
tail -2 gpurun_out/r2_bench_n2_staged.err | cut -c1-300
