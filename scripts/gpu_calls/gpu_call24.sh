#!/bin/bash
# final single-GPU verification of the shipped build: smoke, full -m gpu suite, both bench arms, the other configs
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | grep smoke | tee $O/smoke24.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/pytest_gpu24.log
timeout 900 python bench.py > $O/r02_bench_cfg3_n1.json 2> $O/r02_bench_cfg3_n1.err; tail -c 600 $O/r02_bench_cfg3_n1.json; tail -3 $O/r02_bench_cfg3_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_cfg3_reference_arm.json 2> $O/r02_bench_ref.err; tail -c 300 $O/r02_bench_cfg3_reference_arm.json
for c in cfg1 cfg2 cfg3o cfg4; do timeout 600 python bench.py --config $c --steps 50 --no-extra > $O/r02_bench_${c}_n1.json 2> $O/r02_bench_${c}.err; cut -c1-260 $O/r02_bench_${c}_n1.json; done
