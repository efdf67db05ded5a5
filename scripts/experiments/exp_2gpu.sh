#!/bin/bash
set -u
out=gpurun_out/exp_2gpu
mkdir -p $out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 12 --warmup 3 --no-extras > $out/bench_2gpu.json 2> $out/bench_2gpu.err
tail -1 $out/bench_2gpu.json | cut -c1-400
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/bench_train.py --steps 6 --warmup 3 > $out/train_2gpu.json 2> $out/train_2gpu.err
tail -1 $out/train_2gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 trainer.py --steps 3 > $out/trainer_cli.log 2>&1
tail -3 $out/trainer_cli.log
