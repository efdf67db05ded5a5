#!/bin/bash
# round 2: bench line at N = 8 (AR replicas + data-parallel training leg with the per-slice fused step)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 12 --warmup 3 --no-sweep > gpurun_out/r2_bench_n8_staged.json 2> gpurun_out/r2_bench_n8_staged.err
echo rc=$?
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n8_staged.json').read().strip().splitlines()[-1])
    print('AR', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))
    t=d['train']; print('train', {k:t.get(k) for k in ['ms_per_step','samples_per_s','allreduce_ms','exposed_allreduce_ms','ms_per_step_without_allreduce','ms_per_step_free_running','sync_skew_ms','allreduce_overlap','allreduce_busbw_gbs','error']})
    for k,v in (t.get('modes') or {}).items(): print(k, {a:round(b,3) for a,b in v.items()})
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r2_bench_n8_staged.err').read()[-3000:])
PY
tail -2 gpurun_out/r2_bench_n8_staged.err | cut -c1-300
