#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -k "cfg5_lic_volume_256 or anisotropic or layouts_bit or noise_layouts or lic_volume" 2>&1 | grep -v "^Volume data" | tail -60 > gpurun_out/pytest_gpu6.log
tail -60 gpurun_out/pytest_gpu6.log
