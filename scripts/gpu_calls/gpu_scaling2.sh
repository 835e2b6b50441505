#!/bin/bash
# round 2 multi-GPU record on one box: gpurun --gpus N -- 'bash scripts/gpu_calls/gpu_scaling2.sh N'
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
nvidia-smi --query-gpu=name --format=csv,noheader | head -1; nproc
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 200 --warmup 5 > $O/r02_bench_cfg3_n${N}_p2p.json 2> $O/r02_bench_cfg3_n${N}_p2p.err
cut -c1-330 $O/r02_bench_cfg3_n${N}_p2p.json; grep -o '"extra".*' $O/r02_bench_cfg3_n${N}_p2p.json | cut -c1-500; tail -2 $O/r02_bench_cfg3_n${N}_p2p.err
timeout 400 $TR --master-port 29513 bench.py --gpus $N --steps 100 --warmup 5 --exchange nccl --no-extra > $O/r02_bench_cfg3_n${N}_nccl.json 2> $O/r02_bench_nccl.err
cut -c1-330 $O/r02_bench_cfg3_n${N}_nccl.json
timeout 900 $TR --master-port 29514 bench.py --gpus $N --config cfg5 --steps 2 --warmup 1 > $O/r02_bench_cfg5_n${N}.json 2> $O/r02_bench_cfg5_n${N}.err
cat $O/r02_bench_cfg5_n${N}.json | cut -c1-1600; tail -2 $O/r02_bench_cfg5_n${N}.err
timeout 300 $TR --master-port 29511 scripts/check_dist.py > $O/r02_check_dist_n$N.log 2>&1; echo "check_dist rc=$?"
grep -o "identical True" $O/r02_check_dist_n$N.log | wc -l; grep -o "identical False" $O/r02_check_dist_n$N.log | wc -l
