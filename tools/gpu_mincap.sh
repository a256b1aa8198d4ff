#!/bin/bash
# wide rows: is a small shared-memory cache worth its staging, or should everything stream and the L1 stay large?
qb() { echo "== $*"; env "$@" timeout 300 python tools/quick_bench.py --shape $SHAPE --k $K --implicit $IMP --iters 5 2>&1 | grep -E "RESULT|rror" | cut -c1-90; }
SHAPE=ml10m K=128 IMP=0; qb CMFB200_RES_MINCAP=16; qb CMFB200_RES_MINCAP=1000
SHAPE=lastfm K=128 IMP=1; qb CMFB200_RES_MINCAP=16; qb CMFB200_RES_MINCAP=1000
SHAPE=ml10m K=64 IMP=0; qb CMFB200_RES_MINCAP=1000
SHAPE=lastfm K=64 IMP=1; qb CMFB200_RES_MINCAP=1000
