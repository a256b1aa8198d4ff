#!/bin/bash
mkdir -p gpurun_out/r6
S=$(date +%s)
timeout 300 python tools/e2e_timing.py 2>&1 | tail -30
for w in ml10m_explicit_cg_k64_f32 lastfm_implicit_cg_k64_f32; do
  timeout 900 python bench.py --workload $w > gpurun_out/r6/bench_$w.json 2> gpurun_out/r6/bench_$w.err; cut -c1-200 gpurun_out/r6/bench_$w.json
done
timeout 600 python bench.py --impl reference --steps 10 > gpurun_out/r6/bench_reference.json 2> gpurun_out/r6/bench_reference.err; cut -c1-300 gpurun_out/r6/bench_reference.json
for w in ml10m_explicit_cg_k128_f32 lastfm_implicit_cg_k128_f32 lastfm_implicit_cg_k256_f32 ml10m_explicit_chol_k64_f32; do
  timeout 900 python bench.py --workload $w --steps 5 --no-cpu-baseline > gpurun_out/r6/bench_$w.json 2> gpurun_out/r6/bench_$w.err; cut -c1-200 gpurun_out/r6/bench_$w.json
done
echo "total $(( $(date +%s) - S )) s"
