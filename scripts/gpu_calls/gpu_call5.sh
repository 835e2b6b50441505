#!/bin/bash
# round 2, call 5: A/B of the guard-band / shared-cell walk, the whole GPU suite (incl. whole-frame parity), the bench line
mkdir -p gpurun_out
V=vectorvisualization_b200
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
LIBS="$V/libvv_b200_r1.so $V/libvv_b200.so@WALK_FAST_PATHS:0 $V/libvv_b200.so $V/libvv_b200_regs80.so $V/libvv_b200_t128x7.so"
for c in cfg3 cfg2 cfg1; do timeout 600 python scripts/ab.py cfg=$c loop=20 $LIBS; done 2>&1 | tee gpurun_out/ab5.log
timeout 600 python scripts/ab.py cfg=cfg4 loop=3 $LIBS 2>&1 | tee -a gpurun_out/ab5.log
timeout 1500 python -m pytest tests -m gpu -q --durations=15 2>&1 | tail -60 > gpurun_out/pytest_gpu5.log
tail -25 gpurun_out/pytest_gpu5.log
timeout 900 python bench.py > gpurun_out/bench5_cfg3.json 2> gpurun_out/bench5_cfg3.err
tail -c 3000 gpurun_out/bench5_cfg3.json; tail -5 gpurun_out/bench5_cfg3.err
