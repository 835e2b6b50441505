#!/bin/bash
mkdir -p gpurun_out
V=vectorvisualization_b200
for tag in "" _reuse80; do
  VV_B200_LIB=$PWD/$V/libvv_b200$tag.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:lic_sample -s 1 -c 1 -o gpurun_out/r02_lic_sample_cfg3_reuse${tag:-_64} -f python scripts/profile_frame.py cfg3 2 > gpurun_out/ncu_full$tag.log 2>&1
  tail -2 gpurun_out/ncu_full$tag.log
done
ls -la gpurun_out/*.ncu-rep
