#!/bin/bash
mkdir -p gpurun_out
V=vectorvisualization_b200
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^Volume data" | tail -30 > gpurun_out/pytest_gpu9.log
tail -30 gpurun_out/pytest_gpu9.log
for c in cfg1 cfg2 cfg3 cfg3o; do timeout 600 python scripts/ab.py cfg=$c loop=50 $V/libvv_b200_r1.so $V/libvv_b200.so $V/libvv_b200.so@DEPTH_MAJOR:0; done 2>&1 | tee gpurun_out/ab9.log
