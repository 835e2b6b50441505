#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/frames.log
for lib in libvv_b200.so libvv_b200_c1.so libvv_b200_c2.so; do
  echo "== $lib" >> gpurun_out/frames.log
  for cfg in cfg3 cfg2; do VV_B200_LIB=$PWD/vectorvisualization_b200/$lib python scripts/profile_frame.py $cfg 3 >> gpurun_out/frames.log 2>&1; done
done
grep -E "==|frame 2" gpurun_out/frames.log
