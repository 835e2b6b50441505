#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_cfg3_n8.log 2>&1; tail -1 gpurun_out/bench_cfg3_n8.log | cut -c1-330
timeout 600 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --config cfg4 > gpurun_out/bench_cfg4_n8.log 2>&1; tail -1 gpurun_out/bench_cfg4_n8.log | cut -c1-330
timeout 400 $TR --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_cfg3_n4.log 2>&1; tail -1 gpurun_out/bench_cfg3_n4.log | cut -c1-330
timeout 400 $TR --nproc-per-node 2 --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_cfg3_n2.log 2>&1; tail -1 gpurun_out/bench_cfg3_n2.log | cut -c1-330
timeout 300 $TR --nproc-per-node 8 --master-port 29511 scripts/check_dist.py > gpurun_out/check_dist_n8.log 2>&1; echo "check_dist rc=$?"; grep -o "identical True" gpurun_out/check_dist_n8.log | wc -l
