#!/bin/bash
# shipped build of this session (march checkpoints, split walk loops, 7 x 128 CTAs for the gradient build, xy-quad field by default):
# full GPU suite, both bench arms, launch list, ncu captures of lic_sample on cfg3 / cfg4, sanitizer tools, LIC volume 1024^3 layouts
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/pytest_gpu17.log
timeout 900 python bench.py > $O/r02_bench_cfg3_n1.json 2> $O/r02_bench_cfg3_n1.err; tail -c 1200 $O/r02_bench_cfg3_n1.json; tail -3 $O/r02_bench_cfg3_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_cfg3_reference_arm.json 2> $O/r02_bench_ref.err; tail -c 500 $O/r02_bench_cfg3_reference_arm.json
for c in cfg1 cfg2 cfg3o; do timeout 600 python bench.py --config $c --steps 100 --no-extra > $O/r02_bench_${c}_n1.json 2> $O/r02_bench_${c}.err; cut -c1-330 $O/r02_bench_${c}_n1.json; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r02_launches_cfg3.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $O/ncu_list.log 2>&1
bash scripts/ncu_capture.sh r02b_lic_sample_cfg3 lic_sample 1 1 python scripts/profile_frame.py cfg3 2
bash scripts/ncu_capture.sh r02b_lic_sample_cfg4 lic_sample 1 0 python scripts/profile_frame.py cfg4 2
for tool in racecheck synccheck memcheck; do timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_smoke.py > $O/r02_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -2 $O/r02_sanitizer_$tool.log; done
for lay in 1 2; do timeout 900 python scripts/run_licvol.py 1024 4096 cfg5 $lay 2>&1 | grep -E "lic_volume|sha1" | tail -3; done | tee $O/licvol17.log
