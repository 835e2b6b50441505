#!/bin/bash
mkdir -p gpurun_out
bash scripts/ncu_capture.sh r02b_lic_sample_cfg3_part8 lic_sample 1 0 python scripts/profile_frame.py cfg3 2 part=0/8
