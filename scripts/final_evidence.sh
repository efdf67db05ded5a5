#!/bin/bash
# Round-end evidence on one B200: tests, ncu launch lists, ncu --set full of the hot kernels, bench lines.
set -u
out=gpurun_out/final
mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on -f"
(timeout 1200 python -m pytest tests -q -m gpu > $out/pytest_gpu.log 2>&1; tail -2 $out/pytest_gpu.log)
# --- ncu --set full: inference kernels at the bench shape (bench.py --kernels-only launches each GEMM 11 times)
prof_inf() {  # name kernel-regex skip
  timeout 300 $NCU -k regex:$2 -s $3 -c 1 -o $out/$1 python bench.py --kernels-only > $out/$1.log 2>&1
}
prof_inf gemm_qkv gemm_tc2_kernel 5
prof_inf gemm_out gemm_tc2_kernel 16
prof_inf gemm_ff1 gemm_tc2_kernel 27
prof_inf gemm_ff2 gemm_tc2_kernel 38
prof_inf sdpa sdpa_tc_kernel 5
prof_inf ln_split ln_split_kernel 5
prof_trn() {
  timeout 300 $NCU -k regex:$2 -s $3 -c 1 -o $out/$1 python scripts/bench_train.py --steps 1 --warmup 1 > $out/$1.log 2>&1
}
prof_trn wgrad_ff gemm_wgrad2_kernel 40
prof_trn sdpa_bwd sdpa_bwd_tc_kernel 18
prof_trn ln_bwd ln_bwd_kernel 36
python scripts/ncu_summary.py $out/ncu_full_summary.txt $out/dram_traffic.json gemm_qkv=$out/gemm_qkv.ncu-rep gemm_out=$out/gemm_out.ncu-rep gemm_ff1=$out/gemm_ff1.ncu-rep gemm_ff2=$out/gemm_ff2.ncu-rep sdpa=$out/sdpa.ncu-rep ln_split=$out/ln_split.ncu-rep wgrad_ff=$out/wgrad_ff.ncu-rep sdpa_bwd=$out/sdpa_bwd.ncu-rep ln_bwd=$out/ln_bwd.ncu-rep > /dev/null
cp $out/dram_traffic.json profiles/r1_dram_traffic.json
# --- launch lists
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras > $out/launches_bench.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches_train.csv python scripts/bench_train.py --steps 1 --warmup 1 > $out/launches_train.log 2>&1
# --- bench lines (never under a profiler)
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; tail -c 600 $out/bench.json
timeout 300 python scripts/bench_train.py --steps 6 --warmup 3 > $out/train.json 2>/dev/null; tail -1 $out/train.json
timeout 600 python scripts/sweep_batch.py > $out/sweep.json 2> $out/sweep.err; tail -c 400 $out/sweep.json
rm -f $out/*.log.tmp
ls -la $out | head -40
