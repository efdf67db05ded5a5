#!/bin/bash
set -u
out=gpurun_out/exp_sdpa_bwd
mkdir -p $out
timeout 300 python -m pytest tests/test_backward_gpu.py -x -q -m gpu -k "sdpa" > $out/pytest.log 2>&1
tail -6 $out/pytest.log
if grep -q "failed\|error" $out/pytest.log; then exit 1; fi
for f in 0 1; do
  echo "== sdpa_bwd_tc=$f"
  FACT_FLAGS=sdpa_bwd_tc=$f timeout 300 python scripts/bench_train.py --steps 4 --warmup 2 2>/dev/null | tail -1 | python -c "import json,sys; r=json.loads(sys.stdin.read()); print(r['ms_per_step'], r['loss_last'])"
done
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_backward_gpu.py -x -q -m gpu > $out/pytest_train.log 2>&1
tail -3 $out/pytest_train.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sdpa_bwd -c 100 --csv --log-file $out/launches.csv python scripts/bench_train.py --steps 1 --warmup 1 > $out/ncu.log 2>&1
python scripts/summarize_launches.py $out/launches.csv | head -6
