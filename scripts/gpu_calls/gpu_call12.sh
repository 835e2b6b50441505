#!/bin/bash
mkdir -p gpurun_out
bash scripts/ncu_capture.sh r02_lic_sample_cfg3 lic_sample 1 1 python scripts/profile_frame.py cfg3 2
bash scripts/ncu_capture.sh r02_lic_sample_cfg4 lic_sample 1 0 python scripts/profile_frame.py cfg4 2
bash scripts/ncu_capture.sh r02_lic_volume_1024 lic_volume 0 0 python scripts/profile_frame.py cfg5 1 n=1024 size=1024
