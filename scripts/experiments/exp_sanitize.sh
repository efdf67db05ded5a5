#!/bin/bash
set -u
out=gpurun_out/exp_sanitize
mkdir -p $out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py tests/test_backward_gpu.py -q -m gpu -x -k "pair_tma or wgrad_gemm_pair or sdpa_forward_lse or layernorm_backward or embedding_shape or gemm_tc_inplace" > $out/memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" $out/memcheck.log | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/launches_b1.csv python bench.py --batch 1 --steps 3 --warmup 3 --no-extras > $out/b1.log 2>&1
python scripts/summarize_launches.py $out/launches_b1.csv --between step_inc_kernel
