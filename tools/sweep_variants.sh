#!/bin/bash
for cfg in 0 1 3 6 5; do echo "== CFG64=$cfg"; CMFB200_CFG64=$cfg python tools/quick_bench.py --shape ml10m --k 64 --iters 5 2>&1 | grep RESULT; done
CMFB200_STAGED=1 python tools/quick_bench.py --shape ml10m --k 64 --iters 5 2>&1 | grep RESULT
python tools/quick_bench.py --shape lastfm --k 64 --implicit 1 --iters 5 2>&1 | grep RESULT
