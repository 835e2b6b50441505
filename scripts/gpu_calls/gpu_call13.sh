#!/bin/bash
mkdir -p gpurun_out
V=vectorvisualization_b200
L=$V/libvv_b200.so
timeout 600 python -m pytest tests -m gpu -q -x -k "layouts_bit_identical or noise_layouts" 2>&1 | tail -3
for c in cfg3 cfg2 cfg1; do timeout 600 python scripts/ab.py cfg=$c loop=30 $L $L@FIELD_LAYOUT:2; done 2>&1 | tee gpurun_out/ab13.log
timeout 600 python scripts/ab.py cfg=cfg4 loop=3 $L $L@FIELD_LAYOUT:2 2>&1 | tee -a gpurun_out/ab13.log
for lay in 1 2; do timeout 600 python scripts/run_licvol.py 512 2048 cfg5 $lay 2>&1 | grep -E "lic_volume|sha1|volume_raycast" | tail -4; done | tee -a gpurun_out/ab13.log
