#!/bin/bash
# multi-GPU check on one box: gpurun --gpus N -- 'bash scripts/gpu_scaling.sh N'
#   bench.py with both tile exchanges (peer-to-peer stores / NCCL gather), cfg4, and the bit-identity check
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
for ex in p2p nccl; do
  timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 40 --warmup 5 --exchange $ex > gpurun_out/bench_cfg3_n${N}_$ex.log 2>&1
  tail -1 gpurun_out/bench_cfg3_n${N}_$ex.log | cut -c1-250
done
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --config cfg4 > gpurun_out/bench_cfg4_n${N}_p2p.log 2>&1
tail -1 gpurun_out/bench_cfg4_n${N}_p2p.log | cut -c1-250
timeout 300 $TR --master-port 29511 scripts/check_dist.py > gpurun_out/check_dist_n$N.log 2>&1; echo "check_dist rc=$?"
grep -o "identical True" gpurun_out/check_dist_n$N.log | wc -l; grep -o "identical False" gpurun_out/check_dist_n$N.log | wc -l
