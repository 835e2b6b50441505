#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/frames.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for opt in "noiselayout=1" "noiselayout=0"; do
  echo "== $opt" >> gpurun_out/frames.log
  python scripts/profile_frame.py cfg3 3 $opt >> gpurun_out/frames.log 2>&1
done
python scripts/profile_frame.py cfg2 3 >> gpurun_out/frames.log 2>&1
python scripts/profile_frame.py cfg1 3 >> gpurun_out/frames.log 2>&1
grep -E "==|frame 2" gpurun_out/frames.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lic_sample -s 1 -c 1 -o gpurun_out/prof_lic_sample_cfg3_v6 -f python scripts/profile_frame.py cfg3 2 > gpurun_out/ncu_full.log 2>&1
