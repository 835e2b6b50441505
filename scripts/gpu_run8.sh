#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/frames.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "item_order or modes or cfg1_small or cfg3_small" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for opt in "depthmajor=0" "depthmajor=1 band=2" "depthmajor=1 band=4" "depthmajor=1 band=8" "depthmajor=1 band=16"; do
  echo "== $opt" >> gpurun_out/frames.log
  python scripts/profile_frame.py cfg3 3 $opt >> gpurun_out/frames.log 2>&1
  python scripts/profile_frame.py cfg2 3 $opt >> gpurun_out/frames.log 2>&1
done
grep -E "==|frame 2" gpurun_out/frames.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lic_sample -s 1 -c 1 -o gpurun_out/prof_lic_sample_cfg3_v5 -f python scripts/profile_frame.py cfg3 2 > gpurun_out/ncu_full.log 2>&1
