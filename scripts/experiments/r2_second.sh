#!/bin/bash
# round 2, second GPU pass (2 GPUs): B128 diagnostic, remaining new GPU tests, bench under torchrun at N=2
mkdir -p gpurun_out
timeout 600 python scripts/experiments/diag_b128.py > gpurun_out/r2_diag_b128.log 2>&1
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_model_gpu.py -q -m gpu -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r2_tests_b.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
cat gpurun_out/r2_diag_b128.log
tail -30 gpurun_out/r2_tests_b.log
tail -c 2500 gpurun_out/r2_bench_n2.json
tail -5 gpurun_out/r2_bench_n2.err
