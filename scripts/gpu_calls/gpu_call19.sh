#!/bin/bash
# one rank's share of the 8-GPU partition on one GPU: where does the frame time outside lic_sample go?
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python scripts/profile_frame.py cfg3 3 part=0/8 loop=200 2>&1 | tail -4 | tee $O/part8.log
timeout 300 python scripts/profile_frame.py cfg3 3 part=3/8 loop=200 2>&1 | tail -4 | tee -a $O/part8.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/part8_launches.csv python scripts/profile_frame.py cfg3 3 part=0/8 > /dev/null 2>&1
grep -E "ray_reset|lic_sample|composite|unblock|ray_setup|item_bucket|bucket_scan|checkpoint" $O/part8_launches.csv | awk -F'","' '{print $5, $NF}' | tail -16 | tee -a $O/part8.log
timeout 300 python scripts/profile_frame.py cfg4 2 part=0/8 loop=20 2>&1 | tail -3 | tee -a $O/part8.log
