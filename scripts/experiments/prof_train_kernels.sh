#!/bin/bash
# ncu --set full of the training-step kernels (one launch each, cross-layer shapes)
set -u
out=gpurun_out/prof_train
mkdir -p $out
for k in ln_bwd_kernel gemm_wgrad2_kernel sdpa_bwd_tc_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 1 -o $out/$k -f python scripts/bench_train.py --steps 1 --warmup 1 > $out/$k.log 2>&1
  ncu -i $out/$k.ncu-rep --page raw --csv 2>/dev/null | python scripts/experiments/ncu_pick.py > $out/$k.txt
  echo "== $k"; cat $out/$k.txt
done
