#!/bin/bash
# developer GPU session: parity suite, then timing of the resident kernel against the direct one
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/test.log; cat gpurun_out/r2/test.log
qb() { echo "== $*"; env "$@" timeout 300 python tools/quick_bench.py --shape $SHAPE --k $K --implicit $IMP --iters 5 2>&1 | grep -E "RESULT|finite|Error|error|assert" ; }
SHAPE=ml10m K=64 IMP=0
qb CMFB200_RESIDENT=0
qb CMFB200_RESIDENT=1
qb CMFB200_RESIDENT=1 CMFB200_RES_CLUSTERS=0
qb CMFB200_RESIDENT=1 CMFB200_RES_OVF8=200
SHAPE=lastfm K=64 IMP=1
qb CMFB200_RESIDENT=0
qb CMFB200_RESIDENT=1
qb CMFB200_RESIDENT=1 CMFB200_RES_CLUSTERS=0
SHAPE=ml10m K=128 IMP=0
qb CMFB200_RESIDENT=0
qb CMFB200_RESIDENT=1
# per-kernel times of one resident iteration and of the configs 3 / 4 steps
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2/launches_resident_ml10m.csv \
   python tools/quick_bench.py --shape ml10m --k 64 --iters 1 > /dev/null 2>&1
for w in ml10m_explicit_cg_k64_f32_implicit_features ml10m_explicit_chol_k128_f64_sideinfo; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_$w.csv \
     python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
done
ls -la gpurun_out/r2
