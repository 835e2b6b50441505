#!/bin/bash
# compositing fused into lic_sample_kernel (VV_OPT_FUSE_COMPOSITE, default 1) against the separate composite_kernel
mkdir -p gpurun_out
V=vectorvisualization_b200
L=$V/libvv_b200.so
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu25.log
for c in cfg3 cfg2; do timeout 900 python scripts/ab.py cfg=$c loop=50 $L $L@FUSE_COMPOSITE:0; done 2>&1 | tee gpurun_out/ab25.log
timeout 900 python scripts/ab.py cfg=cfg3 view=close loop=20 $L $L@FUSE_COMPOSITE:0 2>&1 | tee -a gpurun_out/ab25.log
timeout 900 python scripts/ab.py cfg=cfg1t loop=50 $L $L@FUSE_COMPOSITE:0 2>&1 | tee -a gpurun_out/ab25.log
timeout 900 python scripts/ab.py cfg=cfg4 loop=3 $L $L@FUSE_COMPOSITE:0 2>&1 | tee -a gpurun_out/ab25.log
for f in 1 0; do echo "part=0/8 fuse=$f"; timeout 300 python scripts/profile_frame.py cfg3 1 part=0/8 loop=200 fuse=$f 2>&1 | grep loop; done | tee -a gpurun_out/ab25.log
for tool in racecheck memcheck; do timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_smoke.py > gpurun_out/r02_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -1 gpurun_out/r02_sanitizer_$tool.log; done
