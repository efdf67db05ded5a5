#!/bin/bash
# A/B of the pair-GEMM epilogue variants (gemm_tma_store = 0 direct | 1 bulk stores + in-place reduction | 2 bulk stores)
set -u
out=gpurun_out/exp_tma_store
mkdir -p $out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "pair or inplace or big_tiles or gemm_tc" > $out/pytest.log 2>&1
tail -3 $out/pytest.log
for mode in precise bf16; do
  for f in 0 1 2; do
    echo "== $mode flag=$f"
    timeout 300 python bench.py --kernels-only --mode $mode --flag gemm_tma_store=$f 2> $out/kern_${mode}_$f.err | tee $out/kern_${mode}_$f.txt | grep gemm_ | sed 's/algo_tflops.*executed/exec/'
    tail -2 $out/kern_${mode}_$f.err | grep -i "error\|Traceback" 
  done
done
for f in 0 1; do
  echo "== bench flag=$f"; timeout 300 python bench.py --steps 6 --warmup 3 --no-extras --flag gemm_tma_store=$f 2>/dev/null | tail -1 | python -c "import json,sys; r=json.loads(sys.stdin.read()); print(r['value'], r['ms_per_step'], r.get('e2e'), r.get('clocks'))"
  echo "== train flag=$f"; FACT_FLAGS=gemm_tma_store=$f timeout 300 python scripts/bench_train.py --steps 4 --warmup 2 2>/dev/null | tail -1
done
