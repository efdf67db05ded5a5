#!/bin/bash
# Round-end evidence on one B200: tests, ncu launch lists, ncu --set full of the hot kernels, bench lines.
# Everything lands in gpurun_out/final; the files that are judged are copied into profiles/ (tracked) afterwards.
set -u
R=${ROUND_TAG:-r2}
out=gpurun_out/final
mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on -f"
(timeout 1200 python -m pytest tests -q -m gpu > $out/pytest_gpu.log 2>&1; tail -2 $out/pytest_gpu.log)
# --- ncu --set full: inference kernels at the bench shape (bench.py --kernels-only launches every kernel 11 times:
#     gemm_tc2_kernel in the order QKV, out-proj, FF1, FF2, out-proj + LayerNorm warps, FF2 + LayerNorm warps)
prof_inf() {  # name kernel-regex skip
  timeout 300 $NCU -k regex:$2 -s $3 -c 1 -o $out/$1 python bench.py --kernels-only > $out/$1.log 2>&1
}
prof_inf gemm_qkv gemm_tc2_kernel 5
prof_inf gemm_out gemm_tc2_kernel 16
prof_inf gemm_ff1 gemm_tc2_kernel 27
prof_inf gemm_ff2 gemm_tc2_kernel 38
prof_inf gemm_out_ln gemm_tc2_kernel 49
prof_inf gemm_ff2_ln gemm_tc2_kernel 60
prof_inf sdpa sdpa_tc_kernel 5
prof_inf ln_split ln_split_kernel 5
prof_inf wgrad_ff gemm_wgrad2_kernel 5
prof_trn() {
  timeout 300 $NCU -k regex:$2 -s $3 -c 1 -o $out/$1 python scripts/bench_train.py --steps 1 --warmup 1 > $out/$1.log 2>&1
}
prof_trn sdpa_bwd sdpa_bwd_tc2_kernel 18
prof_trn ln_bwd ln_bwd_kernel 36
python scripts/ncu_summary.py $out/ncu_full_summary.txt $out/dram_traffic.json gemm_qkv=$out/gemm_qkv.ncu-rep gemm_out=$out/gemm_out.ncu-rep gemm_ff1=$out/gemm_ff1.ncu-rep gemm_ff2=$out/gemm_ff2.ncu-rep gemm_out_ln=$out/gemm_out_ln.ncu-rep gemm_ff2_ln=$out/gemm_ff2_ln.ncu-rep sdpa=$out/sdpa.ncu-rep ln_split=$out/ln_split.ncu-rep wgrad_ff=$out/wgrad_ff.ncu-rep sdpa_bwd=$out/sdpa_bwd.ncu-rep ln_bwd=$out/ln_bwd.ncu-rep > /dev/null
rm -f $out/gemm_*.ncu-rep $out/ln_*.ncu-rep $out/wgrad_ff.ncu-rep   # the summaries are what is kept (64 MiB merge limit)
# --- launch lists
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras > $out/launches_bench.log 2>&1
python scripts/summarize_launches.py $out/launches_bench.csv --between step_inc_kernel > $out/launches_bench_summary.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/launches_train.csv python scripts/bench_train.py --steps 1 --warmup 1 > $out/launches_train.log 2>&1
python scripts/summarize_launches.py $out/launches_train.csv --between adam_kernel > $out/launches_train_summary.txt 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches_b1.csv python bench.py --batch 1 --steps 4 --warmup 3 --no-extras > $out/launches_b1.log 2>&1
python scripts/summarize_launches.py $out/launches_b1.csv --between step_inc_kernel > $out/launches_b1_summary.txt 2>&1
# --- bench line (never under a profiler): the driver's command
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; tail -c 300 $out/bench.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_reference.json 2> $out/bench_reference.err; tail -c 300 $out/bench_reference.json
ls -la $out | head -40
