#!/bin/bash
# shipped build with the 2, 6, 18 ... window schedule: full GPU suite, smoke, cfg1 / cfg3o bench lines
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | grep smoke | tee $O/smoke24.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/pytest_gpu24.log
for c in cfg1 cfg3o; do timeout 600 python bench.py --config $c --steps 100 --no-extra > $O/r02_bench_${c}_n1.json 2> $O/r02_bench_${c}.err; cut -c1-260 $O/r02_bench_${c}_n1.json; grep -o '"parity".*' $O/r02_bench_${c}_n1.json | cut -c1-300; done
