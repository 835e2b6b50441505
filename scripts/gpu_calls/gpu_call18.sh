#!/bin/bash
# CTA-affine item hand-out (VV_OPT_ITEM_AFFINITY = chunk length) against the global queue
mkdir -p gpurun_out
V=vectorvisualization_b200
L=$V/libvv_b200.so
for c in cfg3 cfg2; do timeout 900 python scripts/ab.py cfg=$c loop=30 $L $L@ITEM_AFFINITY:4 $L@ITEM_AFFINITY:8 $L@ITEM_AFFINITY:16 $L@ITEM_AFFINITY:32; done 2>&1 | tee gpurun_out/ab18.log
timeout 900 python scripts/ab.py cfg=cfg4 loop=3 $L $L@ITEM_AFFINITY:4 $L@ITEM_AFFINITY:8 $L@ITEM_AFFINITY:16 $L@ITEM_AFFINITY:32 2>&1 | tee -a gpurun_out/ab18.log
timeout 900 python scripts/ab.py cfg=cfg1 loop=50 $L $L@ITEM_AFFINITY:8 2>&1 | tee -a gpurun_out/ab18.log
timeout 900 python scripts/ab.py cfg=cfg3 view=close loop=20 $L $L@ITEM_AFFINITY:8 2>&1 | tee -a gpurun_out/ab18.log
