#!/bin/bash
mkdir -p gpurun_out
V=vectorvisualization_b200
LIBS="$V/libvv_b200.so $V/libvv_b200_reuse80.so $V/libvv_b200_t128x7.so $V/libvv_b200_seq64.so $V/libvv_b200_seq56.so $V/libvv_b200_seq48.so"
for c in cfg3 cfg2; do timeout 600 python scripts/ab.py cfg=$c loop=20 $LIBS; done 2>&1 | tee gpurun_out/ab2.log
timeout 600 python scripts/ab.py cfg=cfg4 loop=3 $LIBS 2>&1 | tee -a gpurun_out/ab2.log
# memory-system sensitivity: the same 21.66 M ray samples over smaller (cache-resident) volumes
for n in 32 64 128; do timeout 300 python scripts/profile_frame.py cfg3 1 n=$n size=1024 loop=10 2>&1 | grep loop; done | tee gpurun_out/cfg3_small_volumes.log
