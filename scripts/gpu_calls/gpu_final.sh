#!/bin/bash
# end-of-round verification: parity suite, smoke, bench (both arms), launch list + full ncu capture of the dominant kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
timeout 600 python scripts/run_licvol.py 512 2048 cfg5 > gpurun_out/licvol_512.log 2>&1; grep lic_volume gpurun_out/licvol_512.log | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_cfg3.csv python scripts/profile_frame.py cfg3 2 > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lic_sample -s 1 -c 1 -o gpurun_out/prof_lic_sample_cfg3_final -f python scripts/profile_frame.py cfg3 2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
