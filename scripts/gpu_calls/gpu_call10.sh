#!/bin/bash
mkdir -p gpurun_out
V=vectorvisualization_b200
L=$V/libvv_b200.so
for c in cfg1 cfg1t cfg3o; do timeout 600 python scripts/ab.py cfg=$c loop=50 $V/libvv_b200_r1.so $L $L@WINDOW_GROWTH:100 $L@WINDOW_GROWTH:150 $L@FIRST_WINDOW:32 $L@FIRST_WINDOW:32,WINDOW_GROWTH:100 $L@FIRST_WINDOW:8 $L@DEPTH_MAJOR:0 $L@RAYCAST_MODE:0; done 2>&1 | tee gpurun_out/ab10.log
timeout 600 python -m pytest tests -m gpu -q -x -k "opaque" 2>&1 | tail -3
