#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/frames.log
python scripts/profile_frame.py cfg3 3 loop=20 >> gpurun_out/frames.log 2>&1
python scripts/profile_frame.py cfg3 3 loop=50 part=0/8 >> gpurun_out/frames.log 2>&1
python scripts/profile_frame.py cfg3 3 loop=50 part=3/8 >> gpurun_out/frames.log 2>&1
python scripts/profile_frame.py cfg4 2 loop=5 >> gpurun_out/frames.log 2>&1
python scripts/profile_frame.py cfg4 2 loop=10 part=0/8 >> gpurun_out/frames.log 2>&1
cat gpurun_out/frames.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg3_part8.csv python scripts/profile_frame.py cfg3 2 part=0/8 > gpurun_out/ncu_list.log 2>&1
