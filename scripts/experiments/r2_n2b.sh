#!/bin/bash
# round 2: training leg at N = 2 in every all-reduce mode (same box, back to back), plus the GPU suite with PDL on the
# large-batch GEMMs
mkdir -p gpurun_out
nvidia-smi topo -m | head -4 > gpurun_out/r2_topo_n2.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/r2_tests_full3.log
tail -2 gpurun_out/r2_tests_full3.log
for mode in adam backward none auto; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 8 --warmup 3 --no-sweep --no-cpu --dp-overlap $mode > gpurun_out/r2_bench_n2_$mode.json 2> gpurun_out/r2_bench_n2_$mode.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n2_$mode.json').read().strip().splitlines()[-1])
t=d['train']; print('$mode', 'AR', round(d['value']), {k:t.get(k) for k in ['ms_per_step','allreduce_ms','exposed_allreduce_ms','ms_per_step_without_allreduce','allreduce_overlap','allreduce_calibration_ms','error']})
PY
done
cat gpurun_out/r2_topo_n2.txt
