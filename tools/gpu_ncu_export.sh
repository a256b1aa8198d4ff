#!/bin/bash
# usage: gpu_ncu_export.sh <tag> <kernel regex> <skip> <count> -- <env assignments...> -- <command...>
# captures --set full on the box and brings back only CSV exports (raw metrics + per-instruction source page)
tag=$1; regex=$2; skip=$3; count=$4; shift 5
envs=()
while [ "$1" != "--" ]; do envs+=("$1"); shift; done
shift
mkdir -p gpurun_out/ncu
env "${envs[@]}" timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count -o /tmp/$tag -f "$@" > gpurun_out/ncu/$tag.log 2>&1
ncu -i /tmp/$tag.ncu-rep --page raw --csv > gpurun_out/ncu/$tag.raw.csv 2>/dev/null
ncu -i /tmp/$tag.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu/$tag.src.csv 2>/dev/null
gzip -f gpurun_out/ncu/$tag.src.csv
ncu -i /tmp/$tag.ncu-rep --page source --csv --print-source cuda > gpurun_out/ncu/$tag.cuda.csv 2>/dev/null
gzip -f gpurun_out/ncu/$tag.cuda.csv
ls -la gpurun_out/ncu/
