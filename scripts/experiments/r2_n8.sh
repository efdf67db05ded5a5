#!/bin/bash
# round 2: the bench line under torchrun at N = 8 (AR replicas + batch sweep + data-parallel training leg)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_n8.txt 2>&1
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1])
print('AR', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])
t=d['train']; print('train', {k:t.get(k) for k in ['ms_per_step','samples_per_s','allreduce_ms','exposed_allreduce_ms','ms_per_step_without_allreduce','allreduce_busbw_gbs','allreduce_slices_mb','gpu_topology','clocks','error']})
print([ (p['batch_per_gpu'], round(p['frames_per_s'])) for p in d['batch_sweep']['points']])
PY
tail -5 gpurun_out/r2_bench_n8.err
head -12 gpurun_out/r2_topo_n8.txt
