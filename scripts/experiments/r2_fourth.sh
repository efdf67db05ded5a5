#!/bin/bash
# round 2, 4th GPU pass: full suite with PDL on, then batch-1 A/B (pdl on / off), then launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_tests_pdl.log
tail -4 gpurun_out/r2_tests_pdl.log
for f in 1 0; do
  echo "== pdl=$f"
  FACT_FLAGS=pdl=$f timeout 300 python scripts/sweep_batch.py --modes precise --batches 1,2,4,8 --steps 40 --warmup 5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        r=json.loads(l); print(r['batch'], round(r['frames_per_s'],1), round(r['ms_per_frame_step'],4))
    elif 'rror' in l: print(l)"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_b1_launches_v2.csv python scripts/sweep_batch.py --modes precise --batches 1 --steps 2 --warmup 1 > gpurun_out/r2_b1_ncu_v2.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2_b1_launches_v2.csv --between step_inc_kernel
