#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s -k "full_size" > gpurun_out/pytest_full.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_full.log; grep -E "full-size|passed|failed|rc=|Error|assert" gpurun_out/pytest_full.log | head -20
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
