#!/bin/bash
# developer GPU session: parity suite (timed), then resident vs direct kernel, then the bench line
mkdir -p gpurun_out/r3
S=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40 > gpurun_out/r3/test.log; tail -5 gpurun_out/r3/test.log
echo "tests took $(( $(date +%s) - S )) s"
qb() { echo "== $*"; env "$@" timeout 300 python tools/quick_bench.py --shape $SHAPE --k $K --implicit $IMP --iters 5 2>&1 | grep -E "RESULT|finite|Error|error|assert" ; }
SHAPE=ml10m K=64 IMP=0
qb CMFB200_RESIDENT=0
qb CMFB200_RESIDENT=1
qb CMFB200_RESIDENT=1 CMFB200_RES_MODE=1
SHAPE=lastfm K=64 IMP=1
qb CMFB200_RESIDENT=0
qb CMFB200_RESIDENT=1
qb CMFB200_RESIDENT=1 CMFB200_RES_MODE=1
SHAPE=ml10m K=128 IMP=0
qb CMFB200_RESIDENT=0
qb CMFB200_RESIDENT=1
SHAPE=lastfm K=128 IMP=1
qb CMFB200_RESIDENT=0
qb CMFB200_RESIDENT=1
timeout 600 python bench.py > gpurun_out/r3/bench_default.json 2> gpurun_out/r3/bench_default.err; cat gpurun_out/r3/bench_default.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r3/launches_resident_ml10m.csv \
   python tools/quick_bench.py --shape ml10m --k 64 --iters 1 > /dev/null 2>&1
echo "total $(( $(date +%s) - S )) s"
