#!/bin/bash
# march checkpoints (ray_checkpoint_kernel): parity tests of the shipped build, then A/B against HEAD and register / CTA-shape variants
mkdir -p gpurun_out
V=vectorvisualization_b200
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu14.log
for c in cfg3 cfg2; do timeout 900 python scripts/ab.py cfg=$c loop=30 $V/libvv_b200_head.so $V/libvv_b200.so $V/libvv_b200_ck8.so $V/libvv_b200_macc.so $V/libvv_b200_t128x7.so $V/libvv_b200_t128x5.so $V/libvv_b200_t128x4.so $V/libvv_b200_regs80.so; done 2>&1 | tee gpurun_out/ab14.log
timeout 600 python scripts/ab.py cfg=cfg1 loop=50 $V/libvv_b200_head.so $V/libvv_b200.so 2>&1 | tee -a gpurun_out/ab14.log
timeout 900 python scripts/ab.py cfg=cfg4 loop=3 $V/libvv_b200_head.so $V/libvv_b200.so $V/libvv_b200_t128x7.so $V/libvv_b200_t128x5.so $V/libvv_b200_regs80.so 2>&1 | tee -a gpurun_out/ab14.log
