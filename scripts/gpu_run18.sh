#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python scripts/run_licvol.py 256 1024 cfg2 > gpurun_out/licvol_256.log 2>&1; grep lic_volume gpurun_out/licvol_256.log | tail -1
timeout 900 python scripts/run_licvol.py 512 2048 cfg5 > gpurun_out/licvol_512.log 2>&1; grep lic_volume gpurun_out/licvol_512.log | tail -1
timeout 900 python scripts/run_licvol.py 1024 4096 cfg5 > gpurun_out/licvol_1024.log 2>&1; tail -4 gpurun_out/licvol_1024.log
