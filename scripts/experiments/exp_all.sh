#!/bin/bash
# full GPU check: all gpu tests, inference bench (short), training bench, optional training launch list
set -u
out=gpurun_out/exp_all
mkdir -p $out
timeout 1500 python -m pytest tests -x -q -m gpu > $out/pytest.log 2>&1
tail -4 $out/pytest.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-extras 2>$out/bench.err | tail -1 | tee $out/bench.json | python -c "import json,sys; r=json.loads(sys.stdin.read()); print('bench', r['value'], r['ms_per_step'], r.get('e2e',{}).get('value'), r.get('gpu_launches'), r.get('clocks'))"
timeout 300 python scripts/bench_train.py --steps 4 --warmup 2 2>/dev/null | tail -1 | tee $out/train.json
if [ "${1:-}" = "ncu" ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $out/launches_train.csv python scripts/bench_train.py --steps 1 --warmup 1 > $out/train_ncu.log 2>&1
  python scripts/summarize_launches.py $out/launches_train.csv --between adam_kernel | head -24
fi
