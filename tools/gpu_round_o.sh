#!/bin/bash
mkdir -p gpurun_out/r5
S=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "tests took $(( $(date +%s) - S )) s"
qb() { echo "== $*"; env "$@" timeout 300 python tools/quick_bench.py --shape $SHAPE --k $K --implicit $IMP --iters 5 2>&1 | grep -E "RESULT|Error|error|assert" ; }
SHAPE=ml10m K=128 IMP=0; qb A=1
SHAPE=lastfm K=128 IMP=1; qb A=1
SHAPE=lastfm K=256 IMP=1; qb A=1
SHAPE=ml10m K=32 IMP=0; qb A=1
for w in ml10m_explicit_cg_k64_f32 lastfm_implicit_cg_k64_f32; do
  timeout 900 python bench.py --workload $w > gpurun_out/r5/bench_$w.json 2> gpurun_out/r5/bench_$w.err; cut -c1-400 gpurun_out/r5/bench_$w.json
done
for w in ml10m_explicit_cg_k64_f32_implicit_features ml10m_explicit_chol_k128_f64_sideinfo cfg1_explicit_cg_k16_f64; do
  timeout 900 python bench.py --workload $w --steps 5 --no-cpu-baseline > gpurun_out/r5/bench_$w.json 2> gpurun_out/r5/bench_$w.err; cut -c1-300 gpurun_out/r5/bench_$w.json; tail -2 gpurun_out/r5/bench_$w.err
done
echo "total $(( $(date +%s) - S )) s"
