#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -k "slicing or golden or chain or mc_offset or clip" 2>&1 | grep -v "^Volume data" | tail -70 > gpurun_out/pytest_gpu8.log
tail -70 gpurun_out/pytest_gpu8.log
