#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log
timeout 600 python scripts/run_licvol.py 256 1024 cfg2 > gpurun_out/licvol_256.log 2>&1; tail -7 gpurun_out/licvol_256.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lic_sample -s 1 -c 1 -o gpurun_out/prof_lic_sample_cfg3_r12 -f python scripts/profile_frame.py cfg3 2 > gpurun_out/ncu_full.log 2>&1
