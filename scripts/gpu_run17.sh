#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/frames.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for lib in libvv_b200.so libvv_b200_m5.so; do
  echo "== $lib" >> gpurun_out/frames.log
  for cfg in cfg3 cfg2; do VV_B200_LIB=$PWD/vectorvisualization_b200/$lib python scripts/profile_frame.py $cfg 3 >> gpurun_out/frames.log 2>&1; done
done
for b in 1 2 8 16; do echo "== band $b" >> gpurun_out/frames.log; python scripts/profile_frame.py cfg3 3 band=$b >> gpurun_out/frames.log 2>&1; done
grep -E "==|frame 2" gpurun_out/frames.log
timeout 900 python scripts/run_licvol.py 1024 4096 cfg5 > gpurun_out/licvol_1024.log 2>&1; tail -8 gpurun_out/licvol_1024.log
