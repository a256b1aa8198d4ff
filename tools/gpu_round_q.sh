#!/bin/bash
mkdir -p gpurun_out/r5
S=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_gram.py -m gpu -q -s 2>&1 | tail -5
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "tests took $(( $(date +%s) - S )) s"
qb() { echo "== $*"; env "$@" timeout 300 python tools/quick_bench.py --shape $SHAPE --k $K --implicit $IMP --iters 5 2>&1 | grep -E "RESULT|Error|error|assert" ; }
SHAPE=lastfm K=64 IMP=1; qb CMFB200_GRAM_TC=1; qb CMFB200_GRAM_TC=0
SHAPE=lastfm K=128 IMP=1; qb CMFB200_GRAM_TC=1; qb CMFB200_GRAM_TC=0
SHAPE=lastfm K=256 IMP=1; qb CMFB200_GRAM_TC=1
bash tools/gpu_ncu_export.sh gram_tc_lastfm gram_tc_kernel 2 2 -- A=1 -- python tools/quick_bench.py --shape lastfm --k 64 --implicit 1 --iters 1 > /dev/null 2>&1
echo "total $(( $(date +%s) - S )) s"
