#!/bin/bash
# xy-quad field + xy-quad bf16 noise as the default (VV_OPT_FIELD_LAYOUT = 3), reload-then-eval variant; LIC volume x-pair vs xy-quad
mkdir -p gpurun_out
V=vectorvisualization_b200
L=$V/libvv_b200.so
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu16.log
for c in cfg3 cfg2; do timeout 900 python scripts/ab.py cfg=$c loop=30 $V/libvv_b200_head.so $L $L@FIELD_LAYOUT:1 $V/libvv_b200_rte.so; done 2>&1 | tee gpurun_out/ab16.log
timeout 900 python scripts/ab.py cfg=cfg3 view=close loop=20 $V/libvv_b200_head.so $L $V/libvv_b200_rte.so 2>&1 | tee -a gpurun_out/ab16.log
timeout 900 python scripts/ab.py cfg=cfg1 loop=50 $V/libvv_b200_head.so $L 2>&1 | tee -a gpurun_out/ab16.log
timeout 900 python scripts/ab.py cfg=cfg4 loop=3 $V/libvv_b200_head.so $L $V/libvv_b200_rte.so 2>&1 | tee -a gpurun_out/ab16.log
for lay in 1 2; do timeout 600 python scripts/run_licvol.py 512 2048 cfg5 $lay 2>&1 | grep -E "lic_volume|sha1" | tail -3; done | tee gpurun_out/licvol16.log
for lay in 1 2; do timeout 900 python scripts/run_licvol.py 1024 4096 cfg5 $lay 2>&1 | grep -E "lic_volume|sha1|volume_raycast" | tail -4; done | tee -a gpurun_out/licvol16.log
