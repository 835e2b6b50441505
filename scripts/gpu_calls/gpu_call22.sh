#!/bin/bash
# lic_volume_kernel with an L2 prefetch of the next Heun step's expected cell (libvv_b200_pf.so) against the shipped kernel
mkdir -p gpurun_out
V=vectorvisualization_b200
for lib in libvv_b200.so libvv_b200_pf.so; do echo $lib 512; VV_B200_LIB=$PWD/$V/$lib timeout 600 python scripts/run_licvol.py 512 2048 cfg5 2>&1 | grep -E "lic_volume|sha1" | tail -3; done | tee gpurun_out/licvol22.log
for lib in libvv_b200.so libvv_b200_pf.so; do echo $lib 256 cfg3-like; VV_B200_LIB=$PWD/$V/$lib timeout 600 python scripts/run_licvol.py 256 1024 cfg3 2>&1 | grep -E "lic_volume|sha1" | tail -3; done | tee -a gpurun_out/licvol22.log
for lib in libvv_b200.so libvv_b200_pf.so; do echo $lib 1024; VV_B200_LIB=$PWD/$V/$lib timeout 900 python scripts/run_licvol.py 1024 4096 cfg5 2>&1 | grep -E "lic_volume|sha1" | tail -3; done | tee -a gpurun_out/licvol22.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | grep smoke | tee -a gpurun_out/licvol22.log
