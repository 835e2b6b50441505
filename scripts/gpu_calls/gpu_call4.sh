#!/bin/bash
mkdir -p gpurun_out
V=vectorvisualization_b200
timeout 120 ./scripts/microbench/pipes2 2>&1 | tee gpurun_out/pipes2.log
LIBS="$V/libvv_b200_r1.so $V/libvv_b200.so $V/libvv_b200_reuse80.so $V/libvv_b200_t128x7.so"
timeout 600 python scripts/ab.py cfg=cfg3 loop=20 $LIBS 2>&1 | tee gpurun_out/ab3.log
timeout 600 python -m pytest tests -m gpu -x -q -k "noise_layouts or layouts_bit or raycast_parity or golden" 2>&1 | tail -3
