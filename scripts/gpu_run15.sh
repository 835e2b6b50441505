#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/frames.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
python scripts/profile_frame.py cfg3 1 loop=20 >> gpurun_out/frames.log 2>&1
for p in 0 3 5; do python scripts/profile_frame.py cfg3 1 loop=50 part=$p/8 >> gpurun_out/frames.log 2>&1; done
cat gpurun_out/frames.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg3_part8.csv python scripts/profile_frame.py cfg3 2 part=0/8 > gpurun_out/ncu_list.log 2>&1
grep -E "composite|ray_setup" gpurun_out/launches_cfg3_part8.csv | tail -3
