#!/bin/bash
# both walkers' cell loads issued together (libvv_b200_pl.so, VV_PAIR_LOADS=1) against the shipped kernel
mkdir -p gpurun_out
V=vectorvisualization_b200
L=$V/libvv_b200.so
for c in cfg3 cfg2; do timeout 900 python scripts/ab.py cfg=$c loop=30 $L $V/libvv_b200_pl.so; done 2>&1 | tee gpurun_out/ab23.log
timeout 900 python scripts/ab.py cfg=cfg3 view=close loop=20 $L $V/libvv_b200_pl.so 2>&1 | tee -a gpurun_out/ab23.log
timeout 900 python scripts/ab.py cfg=cfg1 loop=50 $L $V/libvv_b200_pl.so 2>&1 | tee -a gpurun_out/ab23.log
timeout 900 python scripts/ab.py cfg=cfg4 loop=3 $L $V/libvv_b200_pl.so 2>&1 | tee -a gpurun_out/ab23.log
