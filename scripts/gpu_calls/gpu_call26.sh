#!/bin/bash
# partition units (VV_OPT_PARTITION_UNIT = 1, 2, 4 blocks): every rank's share of the 8-GPU partition, one after the other on one GPU
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "partition_invariance or p2p_exchange" 2>&1 | tail -4 | tee $O/pytest_gpu26.log
for u in 1 2 4; do for r in 0 1 2 3 4 5 6 7; do echo -n "cfg3 unit=$u rank=$r: "; timeout 300 python scripts/profile_frame.py cfg3 0 part=$r/8 unit=$u loop=100 2>&1 | grep loop; done; done | tee $O/part26.log
for u in 1 2 4; do for r in 0 3 6; do echo -n "cfg4 unit=$u rank=$r: "; timeout 300 python scripts/profile_frame.py cfg4 0 part=$r/8 unit=$u loop=8 2>&1 | grep loop; done; done | tee -a $O/part26.log
