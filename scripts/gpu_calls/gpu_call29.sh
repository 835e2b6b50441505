#!/bin/bash
# windowed frames: depth-major bucket sort per window against tile-major emission by composite_kernel (fewer launches)
mkdir -p gpurun_out
V=vectorvisualization_b200
L=$V/libvv_b200.so
for c in cfg1 cfg3o; do timeout 900 python scripts/ab.py cfg=$c loop=50 $L $L@DEPTH_MAJOR:0; done 2>&1 | tee gpurun_out/ab29.log
timeout 900 python scripts/ab.py cfg=cfg1 view=close loop=50 $L $L@DEPTH_MAJOR:0 2>&1 | tee -a gpurun_out/ab29.log
timeout 900 python scripts/ab.py cfg=cfg3o view=close loop=50 $L $L@DEPTH_MAJOR:0 2>&1 | tee -a gpurun_out/ab29.log
