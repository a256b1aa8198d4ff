#!/bin/bash
mkdir -p gpurun_out/r4
S=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -30 > gpurun_out/r4/test_all.log; tail -12 gpurun_out/r4/test_all.log
echo "tests took $(( $(date +%s) - S )) s"
qb() { echo "== $*"; env "$@" timeout 300 python tools/quick_bench.py --shape $SHAPE --k $K --implicit $IMP --iters 5 2>&1 | grep -E "RESULT|finite|Error|error|assert" ; }
SHAPE=ml10m K=64 IMP=0
qb CMFB200_PANEL=1
qb CMFB200_PANEL=1 CMFB200_PANEL_MAXCL=8
qb CMFB200_PANEL=1 CMFB200_PANEL_CLUSTERS=0
qb CMFB200_PANEL=0
SHAPE=lastfm K=64 IMP=1
qb CMFB200_PANEL=1
qb CMFB200_PANEL=0
SHAPE=ml10m K=128 IMP=0
qb CMFB200_PANEL=1
SHAPE=ml10m K=32 IMP=0
qb CMFB200_PANEL=1
qb CMFB200_PANEL=0
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r4/launches_panel_ml10m.csv \
   python tools/quick_bench.py --shape ml10m --k 64 --iters 1 > /dev/null 2>&1
echo "total $(( $(date +%s) - S )) s"
