#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_sweeps.py -m gpu -q -k "every_team or long_rows" 2>&1 | tail -15 > gpurun_out/r2/test_b.log; cat gpurun_out/r2/test_b.log
qb() { echo "== $*"; env "$@" timeout 300 python tools/quick_bench.py --shape $SHAPE --k $K --implicit $IMP --iters 5 2>&1 | grep -E "RESULT|Error|error|assert" ; }
SHAPE=ml10m K=64 IMP=0
qb CMFB200_RESIDENT=1
qb CMFB200_RESIDENT=1 CMFB200_RES_CLUSTERS=0
qb CMFB200_RESIDENT=1 CMFB200_RES_OVF8=200
SHAPE=lastfm K=64 IMP=1
qb CMFB200_RESIDENT=1
SHAPE=ml10m K=128 IMP=0
qb CMFB200_RESIDENT=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2/launches_resident_ml10m_b.csv \
   python tools/quick_bench.py --shape ml10m --k 64 --iters 1 > /dev/null 2>&1
