#!/bin/bash
# fast / slow walk loops as separate instantiations + per-build CTA shape: A/B against HEAD and the single-shape builds
mkdir -p gpurun_out
V=vectorvisualization_b200
timeout 900 python -m pytest tests -m gpu -q -x -k "full_size or layouts or random or raycast_parity" 2>&1 | tail -3 | tee gpurun_out/pytest_gpu15.log
for c in cfg3 cfg2; do timeout 900 python scripts/ab.py cfg=$c loop=30 $V/libvv_b200_head.so $V/libvv_b200.so $V/libvv_b200_s256x4.so $V/libvv_b200_s128x7.so $V/libvv_b200.so@FIELD_LAYOUT:2; done 2>&1 | tee gpurun_out/ab15.log
timeout 900 python scripts/ab.py cfg=cfg3 view=close loop=30 $V/libvv_b200_head.so $V/libvv_b200.so 2>&1 | tee -a gpurun_out/ab15.log
timeout 900 python scripts/ab.py cfg=cfg3o loop=30 $V/libvv_b200_head.so $V/libvv_b200.so 2>&1 | tee -a gpurun_out/ab15.log
timeout 900 python scripts/ab.py cfg=cfg4 loop=3 $V/libvv_b200_head.so $V/libvv_b200.so 2>&1 | tee -a gpurun_out/ab15.log
