#!/bin/bash
# band height of the depth-major item order on one rank's share of the 8-GPU partition (and on the whole frame)
mkdir -p gpurun_out
O=gpurun_out
for b in 1 2 4 8 16 64; do echo "part=0/8 band=$b"; timeout 300 python scripts/profile_frame.py cfg3 1 part=0/8 loop=100 band=$b 2>&1 | grep loop; done | tee $O/band20.log
for b in 2 4 8; do echo "whole frame band=$b"; timeout 300 python scripts/profile_frame.py cfg3 1 loop=50 band=$b 2>&1 | grep loop; done | tee -a $O/band20.log
for b in 2 4 8 16; do echo "cfg4 part=0/8 band=$b"; timeout 300 python scripts/profile_frame.py cfg4 1 part=0/8 loop=10 band=$b 2>&1 | grep loop; done | tee -a $O/band20.log
