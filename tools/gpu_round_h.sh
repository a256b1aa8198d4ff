#!/bin/bash
mkdir -p gpurun_out/r3
S=$(date +%s)
qb() { echo "== $*"; env "$@" timeout 300 python tools/quick_bench.py --shape $SHAPE --k $K --implicit $IMP --iters 5 2>&1 | grep -E "RESULT|Error|error|assert" ; }
SHAPE=ml10m K=64 IMP=0
qb CMFB200_RESIDENT=1 CMFB200_RES_MODE=0
qb CMFB200_RESIDENT=1 CMFB200_RES_MODE=0 CMFB200_RES_BPS=1
qb CMFB200_RESIDENT=1 CMFB200_RES_MODE=1 CMFB200_RES_BPS=1
SHAPE=lastfm K=64 IMP=1
qb CMFB200_RESIDENT=1 CMFB200_RES_MODE=0
qb CMFB200_RESIDENT=1 CMFB200_RES_MODE=0 CMFB200_RES_BPS=1
CMFB200_RES_MODE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cg_resident -s 8 -c 4 -o gpurun_out/r3/full_teams_ml10m -f \
   python tools/quick_bench.py --shape ml10m --k 64 --iters 1 > gpurun_out/r3/ncu_full2.log 2>&1
CMFB200_RESIDENT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cg_sweep -s 6 -c 3 -o gpurun_out/r3/full_direct_ml10m -f \
   python tools/quick_bench.py --shape ml10m --k 64 --iters 1 > gpurun_out/r3/ncu_full3.log 2>&1
timeout 900 python tools/diag_implicit2.py 2>&1 | tee gpurun_out/r3/diag_implicit2.log
echo "total $(( $(date +%s) - S )) s"
