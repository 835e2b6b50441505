#!/bin/bash
# round 2, first GPU call: gather microbench, A/B of the cell-reuse / register variants, parity suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/gpu.txt
timeout 120 ./scripts/microbench/gather > gpurun_out/gather.log 2>&1; cat gpurun_out/gather.log
V=vectorvisualization_b200
LIBS="$V/libvv_b200_r1.so $V/libvv_b200_noreuse64.so $V/libvv_b200_noreuse80.so $V/libvv_b200.so $V/libvv_b200_reuse80.so"
for c in cfg3 cfg2 cfg1; do timeout 600 python scripts/ab.py cfg=$c loop=20 $LIBS; done 2>&1 | tee gpurun_out/ab1.log
timeout 600 python scripts/ab.py cfg=cfg4 loop=3 $LIBS 2>&1 | tee -a gpurun_out/ab1.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
