#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python -m pytest tests -m "not gpu" -x -q > gpurun_out/pytest_cpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_cpu.log; tail -3 gpurun_out/pytest_cpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_dist.py > gpurun_out/check_dist.log 2>&1; echo "check_dist rc=$?" >> gpurun_out/check_dist.log; grep -E "rank|rc=" gpurun_out/check_dist.log | tail -8
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.log 2>&1; tail -1 gpurun_out/bench_n2.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
timeout 900 python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_cfg4_n1.log 2>&1; tail -1 gpurun_out/bench_cfg4_n1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config cfg4 --steps 5 --warmup 3 > gpurun_out/bench_cfg4_n2.log 2>&1; tail -1 gpurun_out/bench_cfg4_n2.log
