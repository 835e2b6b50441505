#!/bin/bash
# two-GPU check of the shipped build: bit-identity of distributed frames / LIC volume (NCCL + peer-to-peer), bench line at N = 2
N=2
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
timeout 300 $TR --master-port 29511 scripts/check_dist.py > $O/r02_check_dist_n$N.log 2>&1; echo "check_dist rc=$?"
grep -o "identical True" $O/r02_check_dist_n$N.log | wc -l; grep -o "identical False" $O/r02_check_dist_n$N.log | wc -l
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 200 --warmup 5 > $O/r02_bench_cfg3_n${N}_p2p.json 2> $O/r02_bench_cfg3_n${N}_p2p.err
cut -c1-330 $O/r02_bench_cfg3_n${N}_p2p.json; grep -o '"extra".*' $O/r02_bench_cfg3_n${N}_p2p.json | cut -c1-400; tail -2 $O/r02_bench_cfg3_n${N}_p2p.err
timeout 600 $TR --master-port 29513 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > $O/r02_bench_ref_n$N.json 2>/dev/null; grep -o '"cpu_baseline".*' $O/r02_bench_ref_n$N.json | cut -c1-300
