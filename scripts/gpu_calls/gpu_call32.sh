#!/bin/bash
# the early-termination path under ncu: launch lists of a cfg1 and a cfg3o frame, one full capture of a cfg1 sample-kernel launch
mkdir -p gpurun_out
O=gpurun_out
for c in cfg1 cfg3o; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02b_launches_$c.csv python scripts/profile_frame.py $c 3 > /dev/null 2>&1
  echo "== $c: kernels of the last frame"; grep -E "ray_reset|lic_sample|composite|unblock|item_bucket|bucket_scan" $O/r02b_launches_$c.csv | awk -F'","' '{print $5, $NF}' | tail -12
done | tee $O/windowed32.log
# the third sample-kernel launch of the third frame (a mid-size window)
bash scripts/ncu_capture.sh r02b_lic_sample_cfg1 lic_sample 12 0 python scripts/profile_frame.py cfg1 4
