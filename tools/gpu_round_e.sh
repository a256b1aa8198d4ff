#!/bin/bash
mkdir -p gpurun_out/r2
qb() { echo "== $*"; env "$@" timeout 300 python tools/quick_bench.py --shape $SHAPE --k $K --implicit $IMP --iters 5 2>&1 | grep -E "RESULT|Error|error|assert" ; }
SHAPE=ml10m K=64 IMP=0
qb CMFB200_RESIDENT=1 CMFB200_RES_MODE=1
qb CMFB200_RESIDENT=1 CMFB200_RES_MODE=1 CMFB200_RES_BPS=1
SHAPE=lastfm K=64 IMP=1
qb CMFB200_RESIDENT=1 CMFB200_RES_MODE=1
qb CMFB200_RESIDENT=0
SHAPE=ml10m K=128 IMP=0
qb CMFB200_RESIDENT=1 CMFB200_RES_MODE=1
CMFB200_RESIDENT=1 CMFB200_RES_MODE=1 timeout 600 python -m pytest tests/test_gpu_sweeps.py -m gpu -q -x -k "every_team or long_rows" 2>&1 | tail -3
CMFB200_RESIDENT=1 CMFB200_RES_MODE=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 16 --csv --log-file gpurun_out/r2/launches_resident_mode1_lastfm.csv \
   python tools/quick_bench.py --shape lastfm --k 64 --implicit 1 --iters 1 > /dev/null 2>&1
