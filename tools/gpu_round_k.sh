#!/bin/bash
mkdir -p gpurun_out/r4
S=$(date +%s)
CMFB200_PANEL=0 timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r4/test_all2.log; tail -8 gpurun_out/r4/test_all2.log
echo "tests took $(( $(date +%s) - S )) s"
qb() { echo "== $*"; env "$@" timeout 300 python tools/quick_bench.py --shape $SHAPE --k $K --implicit $IMP --iters 5 2>&1 | grep -E "RESULT|Error|error|assert" ; }
SHAPE=ml10m K=64 IMP=0
qb CMFB200_PANEL=0
qb CMFB200_PANEL=0 CMFB200_RES_BPS=1
qb CMFB200_PANEL=0 CMFB200_RES_MODE=0
SHAPE=lastfm K=64 IMP=1
qb CMFB200_PANEL=0
qb CMFB200_PANEL=0 CMFB200_RES_BPS=1
SHAPE=ml10m K=32 IMP=0
qb CMFB200_PANEL=0
SHAPE=ml10m K=128 IMP=0
qb CMFB200_PANEL=0
bash tools/gpu_ncu_export.sh res2_ml10m cg_resident_kernel 4 2 -- CMFB200_PANEL=0 -- python tools/quick_bench.py --shape ml10m --k 64 --iters 1 > /dev/null 2>&1
echo "total $(( $(date +%s) - S )) s"
