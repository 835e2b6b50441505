#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
for ex in p2p nccl; do
timeout 400 $TR --master-port 29512 bench.py --gpus 8 --steps 40 --warmup 5 --exchange $ex > gpurun_out/bench_cfg3_n8_$ex.log 2>&1; tail -1 gpurun_out/bench_cfg3_n8_$ex.log | cut -c1-250
done
timeout 600 $TR --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --config cfg4 > gpurun_out/bench_cfg4_n8_p2p.log 2>&1; tail -1 gpurun_out/bench_cfg4_n8_p2p.log | cut -c1-250
timeout 300 $TR --master-port 29511 scripts/check_dist.py > gpurun_out/check_dist_n8.log 2>&1; echo "check_dist rc=$?"; grep -o "identical True" gpurun_out/check_dist_n8.log | wc -l; grep -ci "error" gpurun_out/check_dist_n8.log
