#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "p2p" > gpurun_out/pytest_p2p.log 2>&1; tail -2 gpurun_out/pytest_p2p.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
timeout 300 $TR --master-port 29511 scripts/check_dist.py > gpurun_out/check_dist_n2.log 2>&1; echo "check_dist rc=$?"; grep -c "identical True" gpurun_out/check_dist_n2.log
for ex in p2p nccl; do
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --config cfg4 --exchange $ex > gpurun_out/bench_cfg4_n2_$ex.log 2>&1; tail -1 gpurun_out/bench_cfg4_n2_$ex.log | cut -c1-250
done
