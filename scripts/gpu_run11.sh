#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/frames.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log
timeout 600 python scripts/run_licvol.py 256 1024 cfg2 > gpurun_out/licvol_256.log 2>&1; tail -7 gpurun_out/licvol_256.log
timeout 900 python scripts/run_licvol.py 512 2048 cfg5 > gpurun_out/licvol_512.log 2>&1; tail -7 gpurun_out/licvol_512.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_cfg3.csv python scripts/profile_frame.py cfg3 2 > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lic_sample -s 1 -c 1 -o gpurun_out/prof_lic_sample_cfg3_final -f python scripts/profile_frame.py cfg3 2 > gpurun_out/ncu_full.log 2>&1
