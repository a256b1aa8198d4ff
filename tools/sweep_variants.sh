#!/bin/bash
# developer experiment: timing of kernel variants selected through environment switches
for cfg in 0 1 2; do for lr in 1024 2048 8192; do
  echo "== CFG64=$cfg LONG_ROW=$lr"; CMFB200_CFG64=$cfg CMFB200_LONG_ROW=$lr python tools/quick_bench.py --shape ml10m --k 64 --iters 5 2>&1 | grep RESULT
done; done
for occ in 1 2; do echo "== OCC=$occ"; CMFB200_OCC=$occ python tools/quick_bench.py --shape ml10m --k 64 --iters 5 2>&1 | grep RESULT; done
