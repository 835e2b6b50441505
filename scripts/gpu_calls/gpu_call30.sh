#!/bin/bash
# short depth windows emit their items tile-major (composite_kernel), long ones keep the bucket sort: full suite + the windowed frames
mkdir -p gpurun_out
O=gpurun_out
V=vectorvisualization_b200
L=$V/libvv_b200.so
timeout 300 python __graft_entry__.py --smoke 2>&1 | grep smoke | tee $O/smoke24.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/pytest_gpu24.log
for c in cfg1 cfg3o cfg3; do timeout 900 python scripts/ab.py cfg=$c loop=50 $L $L@DEPTH_MAJOR:0; done 2>&1 | tee $O/ab30.log
for c in cfg1 cfg3o; do timeout 600 python bench.py --config $c --steps 100 --no-extra > $O/r02_bench_${c}_n1.json 2> $O/r02_bench_${c}.err; cut -c1-260 $O/r02_bench_${c}_n1.json; done
