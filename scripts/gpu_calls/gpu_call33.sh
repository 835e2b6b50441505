#!/bin/bash
# composite_kernel skips tiles without live rays: full suite, windowed frames
mkdir -p gpurun_out
O=gpurun_out
L=vectorvisualization_b200/libvv_b200.so
timeout 300 python __graft_entry__.py --smoke 2>&1 | grep smoke | tee $O/smoke24.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/pytest_gpu24.log
for c in cfg1 cfg3o cfg3; do timeout 900 python scripts/ab.py cfg=$c loop=50 $L; done 2>&1 | tee $O/ab33.log
timeout 900 python scripts/ab.py cfg=cfg3o view=close loop=50 $L 2>&1 | tee -a $O/ab33.log
