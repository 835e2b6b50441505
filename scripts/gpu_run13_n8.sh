#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
timeout 300 $TR 29511 scripts/check_dist.py > gpurun_out/check_dist_n8.log 2>&1; echo "check_dist rc=$?"; grep -c "identical True" gpurun_out/check_dist_n8.log; tail -3 gpurun_out/check_dist_n8.log
timeout 400 $TR 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_cfg3_n8.log 2>&1; tail -1 gpurun_out/bench_cfg3_n8.log
timeout 600 $TR 29513 bench.py --gpus 8 --steps 10 --warmup 3 --config cfg4 > gpurun_out/bench_cfg4_n8.log 2>&1; tail -1 gpurun_out/bench_cfg4_n8.log
nvidia-smi topo -m > gpurun_out/topo.log 2>&1
