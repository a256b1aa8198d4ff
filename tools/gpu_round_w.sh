#!/bin/bash
S=$(date +%s)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
CMFB200_RES_CFG64=3 timeout 600 python -m pytest tests/test_gpu_sweeps.py -m gpu -q -x -k "every_team or long_rows or half_sweeps" 2>&1 | tail -2
qb() { echo "== $*"; env "$@" timeout 300 python tools/quick_bench.py --shape $SHAPE --k $K --implicit $IMP --iters 5 2>&1 | grep -E "RESULT|Error|error|assert" ; }
SHAPE=ml10m K=64 IMP=0; qb CMFB200_RES_CFG64=1; qb CMFB200_RES_CFG64=3
SHAPE=lastfm K=64 IMP=1; qb CMFB200_RES_CFG64=1; qb CMFB200_RES_CFG64=3
echo "total $(( $(date +%s) - S )) s"
