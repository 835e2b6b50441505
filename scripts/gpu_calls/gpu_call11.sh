#!/bin/bash
# round 2: bench line (both arms), cfg5 on one GPU, launch list + ncu captures (lic_sample cfg3 / cfg4, lic_volume 1024^3), sanitizer tools
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py > $O/r02_bench_cfg3_n1.json 2> $O/r02_bench_cfg3_n1.err; tail -c 1500 $O/r02_bench_cfg3_n1.json; tail -3 $O/r02_bench_cfg3_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_cfg3_reference_arm.json 2> $O/r02_bench_ref.err; tail -c 600 $O/r02_bench_cfg3_reference_arm.json
for c in cfg1 cfg2 cfg3o; do timeout 600 python bench.py --config $c --steps 100 --no-extra > $O/r02_bench_${c}_n1.json 2> $O/r02_bench_${c}.err; cut -c1-400 $O/r02_bench_${c}_n1.json; done
timeout 900 python bench.py --config cfg5 --steps 2 --warmup 1 > $O/r02_bench_cfg5_n1.json 2> $O/r02_bench_cfg5_n1.err; tail -c 1500 $O/r02_bench_cfg5_n1.json; tail -3 $O/r02_bench_cfg5_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r02_launches_cfg3.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $O/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lic_sample -s 1 -c 1 -o $O/r02_lic_sample_cfg3 -f python scripts/profile_frame.py cfg3 2 > $O/ncu_full_cfg3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lic_sample -s 1 -c 1 -o $O/r02_lic_sample_cfg4 -f python scripts/profile_frame.py cfg4 2 > $O/ncu_full_cfg4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lic_sample -s 1 -c 1 -o $O/r02_lic_sample_cfg1 -f python scripts/profile_frame.py cfg1 2 > $O/ncu_full_cfg1.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:lic_volume -c 1 -o $O/r02_lic_volume_1024 -f python scripts/profile_frame.py cfg5 1 n=1024 size=1024 > $O/ncu_full_licvol.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:volume_raycast -c 1 -o $O/r02_volume_raycast_4096 -f python scripts/profile_frame.py cfg5 1 n=256 size=4096 > $O/ncu_full_volray.log 2>&1
ls -la $O/*.ncu-rep | tail -6
for tool in racecheck synccheck memcheck; do timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_smoke.py > $O/r02_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -2 $O/r02_sanitizer_$tool.log; done
