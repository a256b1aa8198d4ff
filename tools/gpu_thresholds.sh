#!/bin/bash
# sweep of the row-length thresholds of the cached CG kernel (block per row / cluster per row)
qb() { echo "== $*"; env "$@" timeout 300 python tools/quick_bench.py --shape $SHAPE --k $K --implicit $IMP --iters 5 2>&1 | grep -E "RESULT|Error|error|assert" | cut -c1-90; }
for SHAPE_IMP in "ml10m 0" "lastfm 1"; do
  set -- $SHAPE_IMP; SHAPE=$1; IMP=$2; K=64
  qb CMFB200_RES_T_BLOCK=1024 CMFB200_RES_T_CLUSTER=8192
  qb CMFB200_RES_T_BLOCK=512 CMFB200_RES_T_CLUSTER=8192
  qb CMFB200_RES_T_BLOCK=2048 CMFB200_RES_T_CLUSTER=8192
  qb CMFB200_RES_T_BLOCK=1024 CMFB200_RES_T_CLUSTER=4096
  qb CMFB200_RES_T_BLOCK=1024 CMFB200_RES_T_CLUSTER=16384
  qb CMFB200_RES_T_BLOCK=4096 CMFB200_RES_T_CLUSTER=16384
done
timeout 600 python bench.py 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('default bench: ms', j['ms_per_step'], 'value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'])"
