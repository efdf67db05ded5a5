#!/bin/bash
# round 2: fused data-parallel optimizer step on 2 GPUs: correctness test, then the training leg in every mode
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_fused_gpu.py tests/test_train_gpu.py -q -m gpu -x -s 2>&1 | tail -25 > gpurun_out/r2_tests_dpfused.log
tail -12 gpurun_out/r2_tests_dpfused.log
for mode in fused none adam; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 8 --warmup 3 --no-sweep --no-cpu --dp-overlap $mode > gpurun_out/r2_bench_n2_$mode.json 2> gpurun_out/r2_bench_n2_$mode.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n2_$mode.json').read().strip().splitlines()[-1])
    t=d['train']; print('$mode', 'AR', round(d['value']), {k:t.get(k) for k in ['ms_per_step','allreduce_ms','exposed_allreduce_ms','ms_per_step_without_allreduce','allreduce_overlap','fused_step','error']})
except Exception as e:
    print('$mode', 'FAILED', e); print(open('gpurun_out/r2_bench_n2_$mode.err').read()[-1500:])
PY
done
