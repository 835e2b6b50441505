#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/frames.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for cfg in cfg1 cfg2 cfg3; do
  python scripts/profile_frame.py $cfg 4 >> gpurun_out/frames.log 2>&1
done
python scripts/profile_frame.py cfg3 3 ctas=2 >> gpurun_out/frames.log 2>&1
python scripts/profile_frame.py cfg3 3 ctas=3 >> gpurun_out/frames.log 2>&1
python scripts/profile_frame.py cfg3 3 camera=close >> gpurun_out/frames.log 2>&1
python scripts/profile_frame.py cfg3 3 layout=0 >> gpurun_out/frames.log 2>&1
cat gpurun_out/frames.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:lic_sample -s 1 -c 1 -o gpurun_out/prof_lic_sample_cfg3_v2 -f python scripts/profile_frame.py cfg3 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
