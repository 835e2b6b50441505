#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for cfg in cfg1 cfg2 cfg3; do
  python scripts/profile_frame.py $cfg 4 >> gpurun_out/frames.log 2>&1
  python scripts/profile_frame.py $cfg 3 mode=0 >> gpurun_out/frames.log 2>&1
  python scripts/profile_frame.py $cfg 3 layout=0 >> gpurun_out/frames.log 2>&1
done
python scripts/profile_frame.py cfg3 3 camera=close >> gpurun_out/frames.log 2>&1
cat gpurun_out/frames.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3.log 2>&1; tail -2 gpurun_out/bench_cfg3.log
# ncu: launch list and one full capture of the dominant kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg3.csv python scripts/profile_frame.py cfg3 2 > gpurun_out/ncu_list.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:lic_sample -s 1 -c 1 -o gpurun_out/prof_lic_sample_cfg3 -f python scripts/profile_frame.py cfg3 2 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:lic_sample -s 1 -c 1 -o gpurun_out/prof_lic_sample_cfg2 -f python scripts/profile_frame.py cfg2 2 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
