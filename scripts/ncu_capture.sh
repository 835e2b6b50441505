#!/bin/bash
# ncu_capture.sh <tag> <kernel regex> <skip> <keep-rep 0|1> <command...>: one `ncu --set full` capture of the first matching launch after
# <skip>, exported as raw CSV (small) under gpurun_out/; the .ncu-rep itself is kept only on request (gpurun copies back <= 64 MiB)
tag=$1; k=$2; skip=$3; keep=$4; shift 4
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o $O/$tag -f "$@" > $O/$tag.log 2>&1
ncu -i $O/$tag.ncu-rep --page raw --csv > $O/$tag.raw.csv 2>/dev/null
ncu -i $O/$tag.ncu-rep --page details --csv > $O/$tag.details.csv 2>/dev/null
[ "$keep" = "1" ] || rm -f $O/$tag.ncu-rep
ls -la $O/$tag.* | awk '{print $5, $9}'
