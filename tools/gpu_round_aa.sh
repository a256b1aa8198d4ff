#!/bin/bash
S=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_sweeps.py tests/test_gpu_fit.py -m gpu -q -x -k "chol or fit" 2>&1 | tail -3
qb() { echo "== $*"; env "$@" timeout 300 python tools/quick_bench.py --shape $SHAPE --k $K --implicit $IMP --solver chol --iters 3 $EXTRA 2>&1 | grep -E "RESULT|Error|error|assert" ; }
SHAPE=ml10m K=64 IMP=0 EXTRA=""; qb A=1
SHAPE=lastfm K=64 IMP=1; qb A=1
SHAPE=ml10m K=128 IMP=0 EXTRA="--dtype f64"; qb A=1
echo "total $(( $(date +%s) - S )) s"
