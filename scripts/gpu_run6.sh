#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/frames.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for c in 8 16 32 64; do
  echo "== chunk $c" >> gpurun_out/frames.log
  python scripts/profile_frame.py cfg3 3 chunk=$c >> gpurun_out/frames.log 2>&1
  python scripts/profile_frame.py cfg2 3 chunk=$c >> gpurun_out/frames.log 2>&1
done
echo "== cfg1, close" >> gpurun_out/frames.log
python scripts/profile_frame.py cfg1 3 >> gpurun_out/frames.log 2>&1
python scripts/profile_frame.py cfg3 3 camera=close >> gpurun_out/frames.log 2>&1
grep -E "==|frame 2" gpurun_out/frames.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lic_sample -s 1 -c 1 -o gpurun_out/prof_lic_sample_cfg3_v3 -f python scripts/profile_frame.py cfg3 2 > gpurun_out/ncu_full.log 2>&1
