#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 4"
timeout 200 $TR --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_cfg3_n4_p2p.log 2>&1; tail -1 gpurun_out/bench_cfg3_n4_p2p.log | cut -c1-240; grep -o '"exchange": "[^"]*"' gpurun_out/bench_cfg3_n4_p2p.log
timeout 200 $TR --master-port 29511 scripts/check_dist.py > gpurun_out/check_dist_n4.log 2>&1; echo "check_dist rc=$?"; grep -o "identical True" gpurun_out/check_dist_n4.log | wc -l; grep -o "identical False" gpurun_out/check_dist_n4.log | wc -l
