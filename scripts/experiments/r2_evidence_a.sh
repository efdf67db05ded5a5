#!/bin/bash
# round 2 (re-entry): baseline evidence of the committed tree on one B200: GPU suite, default bench line, launch lists
set -u
out=gpurun_out/r2a
mkdir -p $out
(timeout 1200 python -m pytest tests -q -m gpu > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log)
date +%s > $out/t0
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$? $(( $(date +%s) - $(cat $out/t0) )) s"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2a/bench.json').read().strip().splitlines()[-1])
    print('AR', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['clocks'])
    print('b1', d.get('batch1')); print('kernels', d.get('kernels'))
    t=d.get('train') or {}; print('train', {k:t.get(k) for k in ['ms_per_step','samples_per_s','algorithmic_tflops_per_gpu','gpu_launches_per_step','roofline','clocks','error']})
    print([(p['batch_per_gpu'], round(p['frames_per_s'])) for p in (d.get('batch_sweep') or {}).get('points', [])])
    print('cpu', d.get('cpu_baseline'))
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r2a/bench.err').read()[-2500:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras > $out/launches_bench.log 2>&1
python scripts/summarize_launches.py $out/launches_bench.csv --between step_inc_kernel > $out/launches_bench_summary.txt 2>&1 || python scripts/summarize_launches.py $out/launches_bench.csv > $out/launches_bench_summary.txt 2>&1
head -16 $out/launches_bench_summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/launches_train.csv python scripts/bench_train.py --steps 1 --warmup 1 > $out/launches_train.log 2>&1
python scripts/summarize_launches.py $out/launches_train.csv --between adam_kernel > $out/launches_train_summary.txt 2>&1
head -24 $out/launches_train_summary.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches_b1.csv python bench.py --batch 1 --steps 4 --warmup 3 --no-extras > $out/launches_b1.log 2>&1
python scripts/summarize_launches.py $out/launches_b1.csv --between step_inc_kernel > $out/launches_b1_summary.txt 2>&1
head -20 $out/launches_b1_summary.txt
