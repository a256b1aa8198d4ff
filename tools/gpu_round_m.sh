#!/bin/bash
S=$(date +%s)
CMFB200_PANEL=0 timeout 600 python -m pytest tests/test_gpu_sweeps.py -m gpu -q -x 2>&1 | tail -3
CMFB200_PANEL=0 CMFB200_RES_CFG64=1 timeout 600 python -m pytest tests/test_gpu_sweeps.py -m gpu -q -x -k "every_team or long_rows or half_sweeps" 2>&1 | tail -3
echo "tests took $(( $(date +%s) - S )) s"
qb() { echo "== $*"; env "$@" timeout 300 python tools/quick_bench.py --shape $SHAPE --k $K --implicit $IMP --iters 5 2>&1 | grep -E "RESULT|Error|error|assert" ; }
SHAPE=ml10m K=64 IMP=0
qb CMFB200_PANEL=0
qb CMFB200_PANEL=0 CMFB200_RES_CFG64=1
qb CMFB200_PANEL=1 CMFB200_PANEL_CLUSTERS=0
SHAPE=lastfm K=64 IMP=1
qb CMFB200_PANEL=0
qb CMFB200_PANEL=0 CMFB200_RES_CFG64=1
SHAPE=ml10m K=128 IMP=0
qb CMFB200_PANEL=0
SHAPE=lastfm K=128 IMP=1
qb CMFB200_PANEL=0
bash tools/gpu_ncu_export.sh res4_ml10m cg_resident_kernel 4 2 -- CMFB200_PANEL=0 CMFB200_RES_CFG64=1 -- python tools/quick_bench.py --shape ml10m --k 64 --iters 1 > /dev/null 2>&1
echo "total $(( $(date +%s) - S )) s"
