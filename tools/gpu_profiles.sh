#!/bin/bash
# round-end evidence: launch list of the default bench command + full captures of the dominant kernels (CSV exports only)
mkdir -p gpurun_out/prof
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/prof/launches_ml10m.csv \
   python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/prof/launches_lastfm.csv \
   python bench.py --workload lastfm_implicit_cg_k64_f32 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
cap() { tag=$1; regex=$2; skip=$3; count=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count -o /tmp/$tag -f "$@" > gpurun_out/prof/$tag.log 2>&1
  ncu -i /tmp/$tag.ncu-rep --page raw --csv > gpurun_out/prof/$tag.raw.csv 2>/dev/null
  ncu -i /tmp/$tag.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/prof/$tag.src.csv.gz; }
cap full_cg_ml10m cg_resident 6 3 python tools/quick_bench.py --shape ml10m --k 64 --iters 1
cap full_cg_lastfm "cg_resident|gram" 10 5 python tools/quick_bench.py --shape lastfm --k 64 --implicit 1 --iters 1
ls -la gpurun_out/prof
