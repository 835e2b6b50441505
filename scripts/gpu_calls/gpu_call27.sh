#!/bin/bash
# depth-window schedules with windows of any length (first window 1..8, growth 200..400 %) on the early-termination frames
mkdir -p gpurun_out
V=vectorvisualization_b200
L=$V/libvv_b200.so
timeout 600 python -m pytest tests -m gpu -q -x -k "whole_frame_cfg1 or whole_frame_cfg3_opaque or raycast_modes or random_scenes" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu27.log
for c in cfg1 cfg3o; do timeout 900 python scripts/ab.py cfg=$c loop=50 $L $L@FIRST_WINDOW:1 $L@FIRST_WINDOW:2 $L@FIRST_WINDOW:4 $L@FIRST_WINDOW:2,WINDOW_GROWTH:300 $L@FIRST_WINDOW:2,WINDOW_GROWTH:400 $L@FIRST_WINDOW:4,WINDOW_GROWTH:300 $L@FIRST_WINDOW:1,WINDOW_GROWTH:400 $L@FIRST_WINDOW:8,WINDOW_GROWTH:300; done 2>&1 | tee gpurun_out/ab27.log
