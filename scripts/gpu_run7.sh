#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/frames.log
for cfg in cfg3 cfg2 cfg1; do python scripts/profile_frame.py $cfg 4 >> gpurun_out/frames.log 2>&1; done
python scripts/profile_frame.py cfg3 3 camera=close >> gpurun_out/frames.log 2>&1
grep -E "frame 3|frame 2" gpurun_out/frames.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lic_sample -s 1 -c 1 -o gpurun_out/prof_lic_sample_cfg3_v4 -f python scripts/profile_frame.py cfg3 2 > gpurun_out/ncu_full.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
