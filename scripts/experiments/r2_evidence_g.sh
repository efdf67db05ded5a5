#!/bin/bash
# round 2, seventh pass: slice-major work order of the weight-gradient GEMM; attention epilogue reorder
set -u
out=gpurun_out/r2g
mkdir -p $out
(timeout 600 python -m pytest tests/test_backward_gpu.py tests/test_kernels_gpu.py -q -m gpu -k "wgrad or sdpa" 2>&1 | tail -4 | cut -c1-300)
for cfg in "wgrad_order=0" "wgrad_order=1"; do
  echo "== kernels-only $cfg"; FACT_FLAGS=$cfg timeout 200 python bench.py --kernels-only 2>&1 | grep -E "^(sdpa|wgrad)" | cut -c1-200
done
for cfg in "wgrad_order=0" "wgrad_order=1" "wgrad_order=0" "wgrad_order=1"; do
  echo "== train $cfg"; FACT_FLAGS=$cfg timeout 300 python scripts/bench_train.py --steps 8 --warmup 3 2>/dev/null | tail -1 | cut -c150-330
done
(timeout 600 python -m pytest tests/test_train_gpu.py -q -m gpu 2>&1 | tail -3 | cut -c1-300)
NCU="ncu --set full --clock-control none -f"
timeout 400 $NCU -k regex:gemm_wgrad2_kernel -s 5 -c 1 -o $out/wgrad python bench.py --kernels-only > $out/wgrad.log 2>&1; tail -1 $out/wgrad.log
ncu -i $out/wgrad.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; vals=rows[2] if len(rows)>2 else rows[1]
for h,v in zip(hdr,vals):
    if h in ('gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum'): print(h,'=',v)
"
