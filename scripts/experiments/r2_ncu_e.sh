#!/bin/bash
# ncu --set full with source counters: the attention core and the pair GEMM with LayerNorm warps (bench.py --kernels-only)
set -u
out=gpurun_out/r2e
mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:sdpa_tc_kernel -s 5 -c 1 -o $out/sdpa python bench.py --kernels-only > $out/sdpa.log 2>&1; tail -2 $out/sdpa.log
timeout 400 $NCU -k regex:gemm_tc2_kernel -s 49 -c 1 -o $out/gemm_out_ln python bench.py --kernels-only > $out/gemm_out_ln.log 2>&1; tail -2 $out/gemm_out_ln.log
timeout 400 $NCU -k regex:gemm_tc2_kernel -s 16 -c 1 -o $out/gemm_out python bench.py --kernels-only > $out/gemm_out.log 2>&1; tail -2 $out/gemm_out.log
ls -la $out
