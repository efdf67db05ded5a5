#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 12 --warmup 3 --no-sweep --no-cpu > gpurun_out/r2_bench_n2_all.json 2> gpurun_out/r2_bench_n2_all.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n2_all.json').read().strip().splitlines()[-1])
    t=d['train']; print('AR', round(d['value']), {k:t.get(k) for k in ['ms_per_step','allreduce_ms','exposed_allreduce_ms','ms_per_step_without_allreduce','ms_per_step_free_running','sync_skew_ms','allreduce_overlap','modes','fused_step','error']})
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r2_bench_n2_all.err').read()[-2500:])
PY
