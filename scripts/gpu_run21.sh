#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/frames.log
for cfg in cfg3 cfg2 cfg1; do python scripts/profile_frame.py $cfg 3 >> gpurun_out/frames.log 2>&1; done
grep -E "frame 2" gpurun_out/frames.log
timeout 600 python scripts/run_licvol.py 256 1024 cfg2 > gpurun_out/licvol_256.log 2>&1; grep lic_volume gpurun_out/licvol_256.log | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
